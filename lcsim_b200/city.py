"""MOSS-shaped city scenarios (BASELINE configs[3], SURVEY.md §8d "C4"): one grid road network, N vehicles with
lane-position homes and road-id routes, i.e. agents the reference drives with LaneIDM (agent.py:167-171).

`make_city` builds the raw network + agents (numpy / python lists), `to_pb` writes it in the reference's scenario
format (map.pb / agents.pb: roads, junction lane groups, `home.lane_position`, `routes[0].driving.road_ids`,
`vehicle_attribute.min_gap`) so the REAL reference consumes the identical bytes (oracle/gen_golden_city.py), and
`pack_city` lowers it to the CSR arrays `lcsim_b200_city_create` takes (include/lcsim_b200.h).

Structure follows what the reference expects of a MOSS map (lcsim/scripts/scenario_converters/moss_convert.py,
lane/lane.py:93-106, road/road.py:90-95, junction/junction.py:90-103, agent/route.py:56-100): roads own 2-3 parallel
lanes (left to right = offset_in_road), junction lanes connect lane k of an in-road to lane k of an out-road and
are grouped per (in_road, out_road); every road carries one BOUNDARY edge 3 m right of its rightmost lane, every
junction the edges of its right-turn lanes.  By default every movement connects ALL lanes of the in-road, so routes
never require a lane change (route.py:148-189 `in_candidate` is always true); `require_lc=True` narrows left turns to
the leftmost lane and picks random end lanes, which makes LaneIDM force lane changes (idm_policy.py:118-167).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from . import packer, proto

LANE_W = 3.2
JUNC_R = 14.0  # half size of the junction box
EDGE_OFFSET = 3.0


def _right(u):
    return np.array([u[1], -u[0]])


def _bezier(p0, p1, p2, n):
    t = np.linspace(0.0, 1.0, n)[:, None]
    return (1 - t) ** 2 * p0 + 2 * (1 - t) * t * p1 + t ** 2 * p2


def _plen(p):
    return float(np.sum(np.linalg.norm(np.diff(p, axis=0), axis=1)))


def make_city(seed: int, grid: int = 3, n_agents: int = 60, n_lanes: int = 2, spacing: float = 120.0,
              max_hops: int = 4, depart_window: float = 20.0, require_lc: bool = False) -> dict:
    rng = np.random.default_rng(seed)
    lanes, roads, juncs = [], [], []
    road_of = {}  # (from junction, to junction) -> road index
    jid = lambda i, j: i * grid + j  # noqa: E731
    jpos = lambda q: np.array([(q // grid) * spacing, (q % grid) * spacing], float)  # noqa: E731
    for q in range(grid * grid):
        i, j = divmod(q, grid)
        for di, dj in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            ni, nj = i + di, j + dj
            if not (0 <= ni < grid and 0 <= nj < grid):
                continue
            a, b = jpos(q), jpos(jid(ni, nj))
            u = (b - a) / np.linalg.norm(b - a)
            r = _right(u)
            rid = len(roads)
            road_of[(q, jid(ni, nj))] = rid
            nl = n_lanes + (1 if rng.random() < 0.3 else 0)
            speed = float(rng.choice([11.18, 13.89, 16.67]))
            lane_ids = []
            for k in range(nl):
                off = (k + 0.5) * LANE_W
                p0, p2 = a + u * JUNC_R + r * off, b - u * JUNC_R + r * off
                mid = (p0 + p2) / 2 + r * rng.uniform(-0.3, 0.3)  # two segments with a slight kink
                lane_ids.append(len(lanes))
                lanes.append(dict(pts=np.stack([p0, mid, p2]), max_speed=speed, width=LANE_W, road=rid, junction=-1,
                                  offset=k, pred=[], succ=[], left=[], right=[]))
            for k, lid in enumerate(lane_ids):
                lanes[lid]["left"] = lane_ids[:k]  # pb order; lane.py:96-98 reverses the left list (nearest first)
                lanes[lid]["right"] = lane_ids[k + 1:]
            edge = lanes[lane_ids[-1]]["pts"] + r * EDGE_OFFSET
            roads.append(dict(id=rid, lane_ids=lane_ids, frm=q, to=jid(ni, nj), u=u, edges=[(edge, proto.ROAD_EDGE_TYPE_BOUNDARY)]))
    ins_of = [[] for _ in range(grid * grid)]
    outs_of = [[] for _ in range(grid * grid)]
    for r in roads:
        ins_of[r["to"]].append(r)
        outs_of[r["frm"]].append(r)
    for q in range(grid * grid):
        ins, outs = ins_of[q], outs_of[q]
        groups, jl, edges = [], [], []
        for ri in ins:
            for ro in outs:
                if ro["to"] == ri["frm"]:
                    continue  # no U-turns
                n_in, n_out = len(ri["lane_ids"]), len(ro["lane_ids"])
                cross = ri["u"][0] * ro["u"][1] - ri["u"][1] * ro["u"][0]
                ks = list(range(min(n_in, n_out)))
                if n_in > len(ks):  # the extra (rightmost) in-lane joins the rightmost out-lane
                    pairs = [(k, k) for k in ks] + [(n_in - 1, n_out - 1)]
                else:
                    pairs = [(k, k) for k in ks]
                if require_lc and cross > 0.5:  # left turns only from the leftmost lane
                    pairs = pairs[:1]
                g = dict(in_road=ri["id"], out_road=ro["id"], lane_ids=[])
                for kin, kout in pairs:
                    lin, lout = ri["lane_ids"][kin], ro["lane_ids"][kout]
                    p0, p2 = lanes[lin]["pts"][-1], lanes[lout]["pts"][0]
                    if abs(cross) < 0.5:
                        ctrl = (p0 + p2) / 2
                    else:  # corner of the two lane lines
                        t = np.dot(p2 - p0, ri["u"])
                        ctrl = p0 + ri["u"] * t
                    lid = len(lanes)
                    lanes.append(dict(pts=_bezier(p0, ctrl, p2, 6), max_speed=8.33, width=LANE_W, road=-1, junction=q,
                                      offset=0, pred=[lin], succ=[lout], left=[], right=[]))
                    lanes[lin]["succ"].append(lid)
                    lanes[lout]["pred"].append(lid)
                    g["lane_ids"].append(lid)
                    jl.append(lid)
                    if cross < -0.5 and kin == n_in - 1:
                        pts = lanes[lid]["pts"]
                        d = np.gradient(pts, axis=0)
                        d /= np.linalg.norm(d, axis=1, keepdims=True)
                        edges.append((pts + np.stack([d[:, 1], -d[:, 0]], axis=1) * EDGE_OFFSET, proto.ROAD_EDGE_TYPE_BOUNDARY))
                groups.append(g)
        juncs.append(dict(id=10000 + q, lane_ids=jl, groups=groups, edges=edges))
    for ln in lanes:
        ln["length"] = _plen(ln["pts"])
    # ---- agents: home slot on a road lane, a random walk of roads, end on the last road (same lane offset)
    taken = set()
    agents = []
    tries = 0
    while len(agents) < n_agents and tries < 50 * n_agents:
        tries += 1
        r0 = int(rng.integers(len(roads)))
        k = int(rng.integers(len(roads[r0]["lane_ids"])))
        lane0 = roads[r0]["lane_ids"][k]
        slot = int(rng.integers(1, max(2, int(lanes[lane0]["length"] // 9.0))))
        if (lane0, slot) in taken:
            continue
        route = [r0]
        cur_k = k
        ok = True
        for _ in range(int(rng.integers(1, max_hops + 1))):
            q = roads[route[-1]]["to"]
            cand = [g for g in juncs[q]["groups"] if g["in_road"] == route[-1]]
            if not require_lc:
                # stay on movements this lane offers (junction lanes keep the offset, or the extra lane merges right)
                cand = [g for g in cand if any(lanes[l]["pred"][0] == roads[route[-1]]["lane_ids"][cur_k] for l in g["lane_ids"])]
            if not cand:
                break
            g = cand[int(rng.integers(len(cand)))]
            if not require_lc:
                jl = [l for l in g["lane_ids"] if lanes[l]["pred"][0] == roads[route[-1]]["lane_ids"][cur_k]][0]
                cur_k = lanes[lanes[jl]["succ"][0]]["offset"]
            route.append(g["out_road"])
        if not ok:
            continue
        end_road = roads[route[-1]]
        end_k = cur_k if not require_lc else int(rng.integers(len(end_road["lane_ids"])))
        end_lane = end_road["lane_ids"][min(end_k, len(end_road["lane_ids"]) - 1)]
        home_s = slot * 9.0 + float(rng.uniform(-1.5, 1.5))
        end_s = float(rng.uniform(0.45, 0.9)) * lanes[end_lane]["length"]
        if len(route) == 1 and end_s - home_s < 20.0:
            continue
        taken.add((lane0, slot))
        dep = 0.0 if rng.random() < 0.6 else round(float(rng.uniform(0, depart_window)), 1)
        agents.append(dict(
            id=5000 + 7 * len(agents) + int(rng.integers(0, 7)), length=float(rng.uniform(3.8, 5.2)),
            width=float(rng.uniform(1.7, 2.1)), max_speed=float(rng.uniform(9.0, 20.0)),
            max_acc=float(rng.uniform(1.5, 3.0)), usual_acc=2.0, max_brake=-float(rng.uniform(3.0, 6.0)),
            usual_brake=-float(rng.uniform(1.5, 3.0)), min_gap=float(rng.uniform(1.5, 3.0)),
            home_lane=lane0, home_s=home_s, end_lane=end_lane, end_s=end_s, route=route, departure_time=dep))
    return dict(seed=seed, lanes=lanes, roads=roads, juncs=juncs, agents=agents, ads_index=0)


def to_pb(raw: dict):
    m = proto.Map()
    m.header.name = f"synthetic-moss-{raw['seed']}"
    for i, ln in enumerate(raw["lanes"]):
        l = m.lanes.add()
        l.id, l.type, l.turn = i, proto.LANE_TYPE_DRIVING, proto.LANE_TURN_STRAIGHT
        l.max_speed, l.length, l.width = ln["max_speed"], ln["length"], ln["width"]
        for p in ln["pts"]:
            n = l.center_line.nodes.add()
            n.x, n.y = float(p[0]), float(p[1])
        for s in ln["succ"]:
            c = l.successors.add(); c.id = s; c.type = 1
        for s in ln["pred"]:
            c = l.predecessors.add(); c.id = s; c.type = 2
        l.left_lane_ids.extend(ln["left"])
        l.right_lane_ids.extend(ln["right"])
        l.parent_id = raw["roads"][ln["road"]]["id"] if ln["road"] >= 0 else raw["juncs"][ln["junction"]]["id"]
        l.center_line_type = proto.CENTER_LINE_TYPE_SURFACE_STREET
    for r in raw["roads"]:
        rp = m.roads.add()
        rp.id = r["id"]
        rp.lane_ids.extend(r["lane_ids"])
        for pts, t in r["edges"]:
            e = rp.road_edges.add()
            e.type = t
            for p in pts:
                n = e.nodes.add(); n.x, n.y = float(p[0]), float(p[1])
    for j in raw["juncs"]:
        jp = m.junctions.add()
        jp.id = j["id"]
        jp.lane_ids.extend(j["lane_ids"])
        for g in j["groups"]:
            gp = jp.driving_lane_groups.add()
            gp.in_road_id, gp.out_road_id = g["in_road"], g["out_road"]
            gp.lane_ids.extend(g["lane_ids"])
        for pts, t in j["edges"]:
            e = jp.road_edges.add()
            e.type = t
            for p in pts:
                n = e.nodes.add(); n.x, n.y = float(p[0]), float(p[1])
    ag = proto.Agents()
    for a in raw["agents"]:
        p = ag.agents.add()
        p.id = a["id"]
        at = p.attribute
        at.type, at.length, at.width = proto.AGENT_TYPE_VEHICLE, a["length"], a["width"]
        at.max_speed, at.max_acceleration, at.usual_acceleration = a["max_speed"], a["max_acc"], a["usual_acc"]
        at.max_braking_acceleration, at.usual_braking_acceleration = a["max_brake"], a["usual_brake"]
        p.vehicle_attribute.min_gap = a["min_gap"]
        p.vehicle_attribute.lane_change_length = 10.0
        p.home.lane_position.lane_id, p.home.lane_position.s = a["home_lane"], a["home_s"]
        sch = p.schedules.add()
        sch.loop_count = 1
        trip = sch.trips.add()
        trip.mode = proto.TRIP_MODE_DRIVE_ONLY
        trip.departure_time = a["departure_time"]
        trip.end.lane_position.lane_id, trip.end.lane_position.s = a["end_lane"], a["end_s"]
        jr = trip.routes.add()
        jr.type = proto.JOURNEY_TYPE_DRIVING
        jr.driving.road_ids.extend(raw["roads"][r]["id"] for r in a["route"])
    ag.ads_index = raw["ads_index"]
    return m, ag


def raw_from_pb(map_pb, agents_pb) -> dict:
    """map.pb / agents.pb of a MOSS-shaped scenario -> the raw structure `pack_city` lowers (dense indices in pb order).
    What Engine.__init__ derives from the same bytes: LaneManager / RoadManager / JunctionManager wiring
    (lane/lane.py:55-106, road/road.py:90-95, junction/junction.py:90-103) and, per agent, Schedule / Route inputs
    (agent/agent.py:106-125, agent/route.py:56-100, agent/schedule.py:51-64)."""
    lane_ix = {l.id: i for i, l in enumerate(map_pb.lanes)}
    road_ix = {r.id: i for i, r in enumerate(map_pb.roads)}
    lanes = []
    for l in map_pb.lanes:
        pts = np.array([[p.x, p.y] for p in l.center_line.nodes], float)
        lanes.append(dict(id=int(l.id), pts=pts, max_speed=float(l.max_speed), width=float(l.width), road=-1, junction=-1,
                          offset=0, pred=[lane_ix[c.id] for c in l.predecessors], succ=[lane_ix[c.id] for c in l.successors],
                          left=[lane_ix[x] for x in l.left_lane_ids], right=[lane_ix[x] for x in l.right_lane_ids],
                          length=_plen(pts)))
    roads = []
    for ri, r in enumerate(map_pb.roads):
        ids = [lane_ix[x] for x in r.lane_ids]
        for k, li in enumerate(ids):
            lanes[li]["road"], lanes[li]["offset"] = ri, k
        roads.append(dict(id=int(r.id), lane_ids=ids,
                          edges=[(np.array([[p.x, p.y] for p in e.nodes], float), int(e.type)) for e in r.road_edges]))
    juncs = []
    for ji, j in enumerate(map_pb.junctions):
        ids = [lane_ix[x] for x in j.lane_ids]
        for li in ids:
            lanes[li]["junction"] = ji
        groups = [dict(in_road=road_ix[g.in_road_id], out_road=road_ix[g.out_road_id], lane_ids=[lane_ix[x] for x in g.lane_ids])
                  for g in j.driving_lane_groups]
        juncs.append(dict(id=int(j.id), lane_ids=ids, groups=groups,
                          edges=[(np.array([[p.x, p.y] for p in e.nodes], float), int(e.type)) for e in j.road_edges]))
    agents = []
    for a in agents_pb.agents:
        if a.home.lane_position.ByteSize() == 0:
            raise NotImplementedError("city scenarios hold lane-position homes only (xy homes: lcsim_b200.packer)")
        if len(a.schedules) != 1 or len(a.schedules[0].trips) != 1 or a.schedules[0].loop_count != 1:
            raise NotImplementedError("lane agents are expected to carry one schedule x one trip, loop_count=1")
        sch, trip = a.schedules[0], a.schedules[0].trips[0]
        if trip.HasField("departure_time"):
            dep = trip.departure_time
        elif sch.HasField("departure_time"):
            dep = sch.departure_time
        elif trip.HasField("wait_time"):
            dep = 0.0 + trip.wait_time
        else:
            dep = 0.0
        if len(trip.routes) == 0:
            raise ValueError(f"agent {a.id}: no route (reference: Route.valid is False)")
        at = a.attribute
        agents.append(dict(
            id=int(a.id), length=float(at.length), width=float(at.width), max_speed=float(at.max_speed),
            max_acc=float(at.max_acceleration), usual_acc=float(at.usual_acceleration),
            max_brake=float(at.max_braking_acceleration), usual_brake=float(at.usual_braking_acceleration),
            min_gap=float(a.vehicle_attribute.min_gap), home_lane=lane_ix[a.home.lane_position.lane_id],
            home_s=float(a.home.lane_position.s), end_lane=lane_ix[trip.end.lane_position.lane_id],
            end_s=float(trip.end.lane_position.s), route=[road_ix[x] for x in trip.routes[0].driving.road_ids],
            departure_time=float(dep)))
    return dict(seed=-1, lanes=lanes, roads=roads, juncs=juncs, agents=agents, ads_index=int(agents_pb.ads_index))


def pack_dir(data_dir: str, **kw) -> "CityStatic":
    return pack_city(raw_from_pb(proto.load_map(os.path.join(data_dir, "map.pb")),
                                 proto.load_agents(os.path.join(data_dir, "agents.pb"))), **kw)


def is_city_dir(data_dir: str) -> bool:
    """True when the scenario's agents have lane-position homes (LaneIDM agents, agent.py:121-131,167-171)."""
    ag = proto.load_agents(os.path.join(data_dir, "agents.pb"))
    return len(ag.agents) > 0 and ag.agents[0].home.lane_position.ByteSize() > 0


def write_city(out_dir: str, raw: dict):
    m, ag = to_pb(raw)
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "map.pb"), "wb") as f:
        f.write(m.SerializeToString())
    with open(os.path.join(out_dir, "agents.pb"), "wb") as f:
        f.write(ag.SerializeToString())


@dataclass
class CityStatic:
    """CSR arrays of one city (include/lcsim_b200.h lcsim_b200_city_static).  Lane / road / junction / agent
    indices are dense positions in pb order; *_pbid map back to the pb ids."""

    start: float
    interval: float
    n_lanes: int
    n_agents: int
    lane_pbid: np.ndarray  # (NL,) i32
    lane_length: np.ndarray  # (NL,) f64  Lane.length (lane.py:64-77: cumulative np.linalg.norm of the raw segments)
    lane_max_speed: np.ndarray  # (NL,) f64
    lane_width: np.ndarray
    lane_road: np.ndarray  # (NL,) i32 parent road index or -1
    lane_junction: np.ndarray  # (NL,) i32 parent junction index or -1
    lane_offset: np.ndarray  # (NL,) i32 offset_in_road
    lane_succ0: np.ndarray  # (NL,) i32 first successor (unique_successor for junction lanes) or -1
    lane_left: np.ndarray  # (NL,) i32 Lane.left_neighbor() or -1 (lane.py:96-98,195-205)
    lane_right: np.ndarray  # (NL,) i32 Lane.right_neighbor() or -1
    lane_node_start: np.ndarray  # (NL+1,) i32
    node_xy: np.ndarray  # (NN,2) f64 raw centre-line nodes
    node_cum: np.ndarray  # (NN,) f64 Lane.line_lengths
    node_dir: np.ndarray  # (NN,) f64 direction of the segment ENDING at the node (Lane.line_directions[i-1]); [first] unused
    lane_eset: np.ndarray  # (NL,) i32 edge set checked by Agent.is_offroad for this lane (road: raw nodes, junction: resampled)
    eset_start: np.ndarray  # (NS+1,) i32
    edge_xy: np.ndarray  # (NE,2)
    edge_dir: np.ndarray  # (NE,2)
    grp_start: np.ndarray  # (NG+1,) i32
    grp_lane: np.ndarray  # (NGL,) i32 junction lane of the group, pb order
    grp_pre_offset: np.ndarray  # (NGL,) i32 offset_in_road of that lane's first predecessor (route.py:93-99)
    agent_id: np.ndarray
    length: np.ndarray
    width: np.ndarray
    idm: np.ndarray  # (N,5) f64 usual_braking_a, max_braking_a, max_a, max_v, min_gap (idm_policy.py:42-51)
    home_lane: np.ndarray
    home_s: np.ndarray
    end_road: np.ndarray  # (N,) i32 road index of the end lane
    end_offset: np.ndarray  # (N,) i32 offset_in_road of the end lane
    end_s: np.ndarray
    depart_tick: np.ndarray
    wait_order: np.ndarray
    route_start: np.ndarray  # (N+1,) i32 into route_road; hop h of agent a uses group route_grp[route_start[a] - a + h]
    route_road: np.ndarray  # road indices of the route
    route_grp: np.ndarray  # (sum(len(route)-1),) i32 group index of every consecutive road pair
    ads_index: int


def pack_city(raw: dict, start: float = 0.0, interval: float = 0.1, max_ticks: int = 100000) -> CityStatic:
    lanes, roads, juncs, agents = raw["lanes"], raw["roads"], raw["juncs"], raw["agents"]
    NL = len(lanes)
    node_start, nodes, cum, ndir = [0], [], [], []
    lane_len = np.zeros(NL)
    for i, ln in enumerate(lanes):
        c = np.asarray(ln["pts"], np.float64)
        length = 0.0
        cum.append(0.0)
        ndir.append(0.0)
        for k in range(1, len(c)):  # lane.py:66-76, same expressions
            length += np.linalg.norm(c[k] - c[k - 1])
            cum.append(float(length))
            ndir.append(float(np.arctan2(c[k][1] - c[k - 1][1], c[k][0] - c[k - 1][0])))
        lane_len[i] = length
        nodes.append(c)
        node_start.append(node_start[-1] + len(c))
    # edge sets: one per road (raw nodes, road.py:351-356) then one per junction (0.5 m-resampled, junction.py:59-78)
    eset_start, exy, edir = [0], [], []
    for r in roads:
        for pts, _ in r["edges"]:
            p = np.asarray(pts, np.float64)
            exy.append(p)
            edir.append(packer.polyline_dir(p))
        eset_start.append(sum(len(x) for x in exy))
    for j in juncs:
        for pts, _ in j["edges"]:
            p = packer.resample_polyline(np.asarray(pts, np.float64))
            exy.append(p)
            edir.append(packer.polyline_dir(p))
        eset_start.append(sum(len(x) for x in exy))
    lane_eset = np.asarray([ln["road"] if ln["road"] >= 0 else len(roads) + ln["junction"] for ln in lanes], np.int32)
    grp_index, grp_start, grp_lane, grp_pre = {}, [0], [], []
    for j in juncs:
        for g in j["groups"]:
            grp_index[(g["in_road"], g["out_road"])] = len(grp_start) - 1
            for l in g["lane_ids"]:
                grp_lane.append(l)
                grp_pre.append(lanes[lanes[l]["pred"][0]]["offset"])
            grp_start.append(len(grp_lane))
    N = len(agents)
    route_start, route_road, route_grp = [0], [], []
    for a in agents:
        route_road.extend(a["route"])
        route_grp.extend(grp_index[(a["route"][h], a["route"][h + 1])] for h in range(len(a["route"]) - 1))
        route_start.append(len(route_road))
    dep = np.asarray([a["departure_time"] for a in agents], np.float64)
    dtick = np.asarray([packer.depart_tick(float(t), start, interval, max_ticks) for t in dep], np.int32)
    order = np.argsort(dep, kind="stable").astype(np.int32)  # manager.py:33-35 sorted() is stable
    i32 = lambda x: np.asarray(x, np.int32)  # noqa: E731
    return CityStatic(
        start=start, interval=interval, n_lanes=NL, n_agents=N,
        lane_pbid=i32([l.get("id", i) for i, l in enumerate(lanes)]), lane_length=lane_len, lane_max_speed=np.asarray([l["max_speed"] for l in lanes], float),
        lane_width=np.asarray([l["width"] for l in lanes], float), lane_road=i32([l["road"] for l in lanes]),
        lane_junction=i32([l["junction"] for l in lanes]), lane_offset=i32([l["offset"] for l in lanes]),
        lane_succ0=i32([l["succ"][0] if l["succ"] else -1 for l in lanes]),
        lane_left=i32([l["left"][-1] if l["left"] else -1 for l in lanes]),  # pb lists leftmost first; the reference reverses
        lane_right=i32([l["right"][0] if l["right"] else -1 for l in lanes]), lane_node_start=i32(node_start),
        node_xy=np.concatenate(nodes), node_cum=np.asarray(cum), node_dir=np.asarray(ndir), lane_eset=lane_eset,
        eset_start=i32(eset_start), edge_xy=np.concatenate(exy) if exy else np.zeros((0, 2)),
        edge_dir=np.concatenate(edir) if edir else np.zeros((0, 2)), grp_start=i32(grp_start), grp_lane=i32(grp_lane),
        grp_pre_offset=i32(grp_pre), agent_id=i32([a["id"] for a in agents]),
        length=np.asarray([a["length"] for a in agents], float), width=np.asarray([a["width"] for a in agents], float),
        idm=np.asarray([[a["usual_brake"], a["max_brake"], a["max_acc"], a["max_speed"], a["min_gap"]] for a in agents], float).reshape(N, 5),
        home_lane=i32([a["home_lane"] for a in agents]), home_s=np.asarray([a["home_s"] for a in agents], float),
        end_road=i32([lanes[a["end_lane"]]["road"] for a in agents]), end_offset=i32([lanes[a["end_lane"]]["offset"] for a in agents]),
        end_s=np.asarray([a["end_s"] for a in agents], float), depart_tick=dtick, wait_order=order,
        route_start=i32(route_start), route_road=i32(route_road), route_grp=i32(route_grp), ads_index=int(raw["ads_index"]))
