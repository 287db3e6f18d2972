"""Synthetic WOMD-shaped scenario generator.

Emits the same structure ``lcsim/scripts/scenario_converters/waymo_convert.py``
writes for a real WOMD scenario (reference: waymo_convert.py:36-38,46,73-111,
123-172): zero roads, ONE junction (id 300000000) that owns every lane, the road
lines and the road edges; agents with ``home.xy_position``, a single schedule
(``loop_count=1``, ``departure_time = k0 * 0.1``) holding one trip whose
``agent_states`` are the valid logged states.  Every coordinate, velocity,
heading and dimension is drawn, rounded to fp32 and stored as that exact double,
so "same inputs" means bit-identical values for the reference engine (which
reads the .pb) and for this one.

Layout: a 3x3 street grid in a 200 m x 200 m box (two lanes per direction),
curbs around the sixteen blocks as road edges (~4k points after the 0.5 m
resampling the engine applies), lane dividers as road lines, and 128 agents
(vehicles following lane paths with a smooth speed profile, ~half of them
static, plus pedestrians and cyclists on the sidewalks).

The generator returns a plain "raw scenario" dict of numpy arrays;
:func:`to_pb` turns it into ``(map_pb, agents_pb)`` messages and
``packer.pack_raw`` into the engine's SoA layout, so large synthetic batches do
not have to round-trip through protobuf.
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np

from . import proto

JUNCTION_ID = 300000000
BOX = 100.0
LANE_W = 3.5
ROAD_HALF = 2 * LANE_W + 0.5  # two lanes per direction + curb offset
CORNER_R = 4.0
CORE = ROAD_HALF + CORNER_R  # half size of an intersection core
DT = 0.1


def _f32(a):
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)


def _line(p0, p1, spacing):
    p0, p1 = np.asarray(p0, float), np.asarray(p1, float)
    n = max(2, int(np.ceil(np.linalg.norm(p1 - p0) / spacing)) + 1)
    t = np.linspace(0.0, 1.0, n)[:, None]
    return p0 + (p1 - p0) * t


def _arc(center, radius, a0, a1, spacing):
    n = max(3, int(np.ceil(abs(a1 - a0) * radius / spacing)) + 1)
    a = np.linspace(a0, a1, n)
    return np.stack([center[0] + radius * np.cos(a), center[1] + radius * np.sin(a)], axis=1)


def _bezier(p0, d0, p1, d1, spacing):
    """Cubic connector leaving p0 along d0 and arriving at p1 along d1."""
    p0, p1 = np.asarray(p0, float), np.asarray(p1, float)
    k = 0.4 * np.linalg.norm(p1 - p0)
    c0, c1 = p0 + k * np.asarray(d0), p1 - k * np.asarray(d1)
    n = max(4, int(np.ceil(1.2 * np.linalg.norm(p1 - p0) / spacing)) + 1)
    t = np.linspace(0.0, 1.0, n)[:, None]
    return ((1 - t) ** 3) * p0 + 3 * ((1 - t) ** 2) * t * c0 + 3 * (1 - t) * t**2 * c1 + t**3 * p1


_DIRS = {"E": (1.0, 0.0), "W": (-1.0, 0.0), "N": (0.0, 1.0), "S": (0.0, -1.0)}
_RIGHT_OF = {"E": "S", "S": "W", "W": "N", "N": "E"}
_LEFT_OF = {"E": "N", "N": "W", "W": "S", "S": "E"}


def _build_network(rng: np.random.Generator):
    """Street grid: returns lanes (list of dict), edges, lines."""
    xs = np.array([-60.0, 0.0, 60.0]) + rng.uniform(-6, 6, 3)
    ys = np.array([-60.0, 0.0, 60.0]) + rng.uniform(-6, 6, 3)
    lanes: List[dict] = []
    seg_index: Dict[tuple, int] = {}

    def lane_offset(l):  # l=0 left (inner) lane, l=1 right (outer) lane, offset to the right of travel
        return LANE_W * (0.5 + l)

    # straight segments between intersection cores
    for orient, centers, cross in (("H", ys, xs), ("V", xs, ys)):
        bounds = [-BOX] + [c for x in cross for c in (x - CORE, x + CORE)] + [BOX]
        for r, c in enumerate(centers):
            for heading in (("E", "W") if orient == "H" else ("N", "S")):
                d = np.array(_DIRS[heading])
                right = np.array([d[1], -d[0]])
                forward = d[0] + d[1] > 0
                for l in range(2):
                    lateral = c + (right[1] if orient == "H" else right[0]) * lane_offset(l)
                    for k in range(4):
                        a, b = bounds[2 * k], bounds[2 * k + 1]
                        if not forward:
                            a, b = b, a
                        p0, p1 = ((a, lateral), (b, lateral)) if orient == "H" else ((lateral, a), (lateral, b))
                        pts = _line(p0, p1, rng.uniform(0.5, 2.0))
                        kk = k if forward else 3 - k  # segment index along the travel direction
                        seg_index[(orient, r, heading, l, kk)] = len(lanes)
                        lanes.append(
                            dict(pts=pts, turn=proto.LANE_TURN_STRAIGHT, succ=[], pred=[],
                                 max_speed=15.65 if l == 0 else 11.18, heading=heading)
                        )

    def connect(a, b):
        lanes[a]["succ"].append(b)
        lanes[b]["pred"].append(a)

    # connectors inside each intersection core
    for i, x in enumerate(xs):
        for j, y in enumerate(ys):
            for heading in "ENWS":
                orient_in = "H" if heading in "EW" else "V"
                r_in = j if orient_in == "H" else i
                # which segment arrives at this core
                k_in = (i if heading == "E" else 2 - i) if orient_in == "H" else (j if heading == "N" else 2 - j)
                for l in range(2):
                    src = seg_index[(orient_in, r_in, heading, l, k_in)]
                    p_in = lanes[src]["pts"][-1]
                    d_in = np.array(_DIRS[heading])
                    moves = [(heading, l, proto.LANE_TURN_STRAIGHT)]
                    if l == 1:
                        moves.append((_RIGHT_OF[heading], 1, 3))
                    else:
                        moves.append((_LEFT_OF[heading], 0, 2))
                    for out_heading, l_out, turn in moves:
                        orient_out = "H" if out_heading in "EW" else "V"
                        r_out = j if orient_out == "H" else i
                        k_out = ((i if out_heading == "E" else 2 - i) if orient_out == "H"
                                 else (j if out_heading == "N" else 2 - j)) + 1
                        dst = seg_index[(orient_out, r_out, out_heading, l_out, k_out)]
                        p_out = lanes[dst]["pts"][0]
                        d_out = np.array(_DIRS[out_heading])
                        spacing = rng.uniform(0.5, 1.5)
                        if turn == proto.LANE_TURN_STRAIGHT:
                            pts = _line(p_in, p_out, spacing)
                        else:
                            pts = _bezier(p_in, d_in, p_out, d_out, spacing)
                        idx = len(lanes)
                        lanes.append(dict(pts=pts, turn=turn, succ=[], pred=[],
                                          max_speed=11.18, heading=out_heading))
                        connect(src, idx)
                        connect(idx, dst)

    # road edges: curb outline of every block, clockwise (road on the left of travel)
    edges: List[Tuple[np.ndarray, int]] = []
    bx = [-BOX] + list(xs) + [BOX]
    by = [-BOX] + list(ys) + [BOX]
    for i in range(4):
        for j in range(4):
            x0 = bx[i] + (ROAD_HALF if i > 0 else 0.0)
            x1 = bx[i + 1] - (ROAD_HALF if i < 3 else 0.0)
            y0 = by[j] + (ROAD_HALF if j > 0 else 0.0)
            y1 = by[j + 1] - (ROAD_HALF if j < 3 else 0.0)
            sp = rng.uniform(0.7, 1.8)
            r = CORNER_R
            # sides present only where a road borders the block
            has = dict(W=i > 0, E=i < 3, S=j > 0, N=j < 3)
            pieces = []
            # clockwise: west side going north, north side going east, east side going south, south going west
            if has["W"]:
                pieces.append(_line((x0, y0 + (r if has["S"] else 0)), (x0, y1 - (r if has["N"] else 0)), sp))
            if has["W"] and has["N"]:
                pieces.append(_arc((x0 + r, y1 - r), r, np.pi, np.pi / 2, sp))
            elif pieces and not has["N"]:
                edges.append((np.concatenate(pieces), proto.ROAD_EDGE_TYPE_BOUNDARY)); pieces = []
            if has["N"]:
                pieces.append(_line((x0 + (r if has["W"] else 0), y1), (x1 - (r if has["E"] else 0), y1), sp))
            if has["N"] and has["E"]:
                pieces.append(_arc((x1 - r, y1 - r), r, np.pi / 2, 0.0, sp))
            elif pieces and not has["E"]:
                edges.append((np.concatenate(pieces), proto.ROAD_EDGE_TYPE_BOUNDARY)); pieces = []
            if has["E"]:
                pieces.append(_line((x1, y1 - (r if has["N"] else 0)), (x1, y0 + (r if has["S"] else 0)), sp))
            if has["E"] and has["S"]:
                pieces.append(_arc((x1 - r, y0 + r), r, 0.0, -np.pi / 2, sp))
            elif pieces and not has["S"]:
                edges.append((np.concatenate(pieces), proto.ROAD_EDGE_TYPE_BOUNDARY)); pieces = []
            if has["S"]:
                pieces.append(_line((x1 - (r if has["E"] else 0), y0), (x0 + (r if has["W"] else 0), y0), sp))
            if has["S"] and has["W"]:
                pieces.append(_arc((x0 + r, y0 + r), r, -np.pi / 2, -np.pi, sp))
            if pieces:
                edges.append((np.concatenate(pieces), proto.ROAD_EDGE_TYPE_BOUNDARY))
    # a few medians on the middle roads (type 2 edges count in the off-road test too)
    for k in range(4):
        a = [-BOX, xs[0] + CORE, xs[1] + CORE, xs[2] + CORE][k]
        b = [xs[0] - CORE, xs[1] - CORE, xs[2] - CORE, BOX][k]
        edges.append((_line((a, ys[1]), (b, ys[1]), rng.uniform(0.8, 1.6)), proto.ROAD_EDGE_TYPE_MEDIAN))
    # tiny degenerate edge (shorter than the 0.5 m resampling interval -> one point, zero direction)
    edges.append((np.array([[bx[1] + 20.0, by[1] + 20.0], [bx[1] + 20.2, by[1] + 20.1]]), proto.ROAD_EDGE_TYPE_BOUNDARY))

    # road lines: lane dividers on every straight segment
    lines: List[Tuple[np.ndarray, int]] = []
    for (orient, r, heading, l, kk), idx in seg_index.items():
        if l != 0:
            continue
        pts = lanes[idx]["pts"]
        d = np.array(_DIRS[heading])
        right = np.array([d[1], -d[0]])
        lines.append((pts[:: max(1, len(pts) // 24)] + right * (LANE_W / 2), 1))  # broken white between lanes
        if heading in "EN":
            lines.append((pts[:: max(1, len(pts) // 24)] - right * (LANE_W / 2), 7))  # double yellow centre
    return lanes, edges, lines


def _path_from(lanes, start, rng, min_len):
    """Random walk on the lane graph; returns a dense polyline of at least min_len metres."""
    pts = [lanes[start]["pts"]]
    cur = start
    total = _length(pts[0])
    while total < min_len and lanes[cur]["succ"]:
        cur = int(rng.choice(lanes[cur]["succ"]))
        pts.append(lanes[cur]["pts"][1:])
        total += _length(lanes[cur]["pts"])
    return np.concatenate(pts)


def _length(p):
    return float(np.sum(np.linalg.norm(np.diff(p, axis=0), axis=1)))


def _sample_path(path, s):
    seg = np.linalg.norm(np.diff(path, axis=0), axis=1)
    cum = np.concatenate([[0.0], np.cumsum(seg)])
    s = np.clip(s, 0.0, cum[-1] - 1e-6)
    i = np.clip(np.searchsorted(cum, s, side="right") - 1, 0, len(seg) - 1)
    f = (s - cum[i]) / np.maximum(seg[i], 1e-9)
    xy = path[i] + (path[i + 1] - path[i]) * f[:, None]
    d = path[i + 1] - path[i]
    h = np.arctan2(d[:, 1], d[:, 0])
    return xy, h


def make_raw_scenario(seed: int, n_agents: int = 128, t_max: int = 91) -> dict:
    """One WOMD-shaped scenario as numpy arrays (fp32-exact doubles)."""
    rng = np.random.default_rng(seed)
    lanes, edges, lines = _build_network(rng)
    for ln in lanes:
        ln["pts"] = _f32(ln["pts"])
    edges = [(_f32(p), t) for p, t in edges]
    lines = [(_f32(p), t) for p, t in lines]

    straight = [i for i, ln in enumerate(lanes) if _length(ln["pts"]) > 30.0]
    agents = []
    for a in range(n_agents):
        is_ads = a == n_agents - 1
        u = rng.random()
        atype = proto.AGENT_TYPE_VEHICLE if (u < 0.85 or is_ads) else (
            proto.AGENT_TYPE_PEDESTRIAN if u < 0.97 else proto.AGENT_TYPE_CYCLIST)
        if atype == proto.AGENT_TYPE_VEHICLE:
            length, width = rng.uniform(3.8, 5.5), rng.uniform(1.7, 2.2)
        elif atype == proto.AGENT_TYPE_PEDESTRIAN:
            length, width = rng.uniform(0.6, 1.0), rng.uniform(0.6, 1.0)
        else:
            length, width = rng.uniform(1.6, 2.0), rng.uniform(0.6, 0.9)
        if is_ads:
            k0, n = 0, t_max
        else:
            k0 = 0 if rng.random() < 0.55 else int(rng.integers(0, t_max - 1))
            n = (t_max - k0) if rng.random() < 0.35 else int(rng.integers(1, t_max - k0 + 1))
        static = (rng.random() < 0.5) and not is_ads
        # speed profile
        t = np.arange(n) * DT
        if static:
            v = np.zeros(n)
        else:
            v0 = abs(rng.normal(6.0, 4.0)) if atype == proto.AGENT_TYPE_VEHICLE else rng.uniform(0.5, 2.5)
            v0 = min(v0, 24.0)
            acc = rng.normal(0.0, 0.8)
            v = np.clip(v0 + acc * t + 0.8 * np.sin(t * rng.uniform(0.3, 1.2) + rng.uniform(0, 6.28)), 0.0, 24.0)
            if rng.random() < 0.2:  # brake to a stop and wait
                stop = int(rng.integers(n // 3, n)) if n > 3 else n
                v[stop:] = 0.0
        s = np.concatenate([[0.0], np.cumsum(v[:-1] * DT)]) if n > 1 else np.zeros(1)
        start = int(rng.choice(straight))
        path = _path_from(lanes, start, rng, s[-1] + 30.0)
        s0 = rng.uniform(0.0, max(1.0, min(_length(lanes[start]["pts"]) - 1.0, 40.0)))
        xy, h = _sample_path(path, s + s0)
        if atype != proto.AGENT_TYPE_VEHICLE:
            # walk on the sidewalk, outside the curb
            off = ROAD_HALF + rng.uniform(0.5, 2.5)
            xy = xy + np.stack([np.sin(h), -np.cos(h)], axis=1) * (off - LANE_W * 0.5)
        elif static and rng.random() < 0.5:
            # parked at the curb
            xy = xy + np.stack([np.sin(h), -np.cos(h)], axis=1) * rng.uniform(1.0, 2.5)
        if static:
            if rng.random() < 0.5:  # sensor jitter around a fixed pose; else exact duplicates (tie cases)
                xy = xy + rng.normal(0, 0.01, xy.shape)
                h = h + rng.normal(0, 0.003, h.shape)
            vel = rng.normal(0, 0.01, (n, 2)) * (rng.random() < 0.5)
        else:
            xy = xy + rng.normal(0, 0.01, xy.shape)
            h = np.unwrap(h) + rng.normal(0, 0.004, h.shape) + (2 * np.pi if rng.random() < 0.1 else 0.0)
            vel = np.stack([v * np.cos(h), v * np.sin(h)], axis=1) + rng.normal(0, 0.02, (n, 2))
        agents.append(dict(
            id=1000 + 3 * a + int(rng.integers(0, 3)), type=atype,
            length=float(_f32(length)), width=float(_f32(width)),
            k0=k0, xy=_f32(xy), vel=_f32(vel), heading=_f32(h),
        ))
    return dict(seed=seed, lanes=lanes, edges=edges, lines=lines, agents=agents, ads_index=n_agents - 1)


def to_pb(raw: dict):
    """Raw scenario -> (map_pb, agents_pb) in the converter's structure."""
    m = proto.Map()
    m.header.name = f"synthetic-womd-{raw['seed']}"
    junc = m.junctions.add()
    junc.id = JUNCTION_ID
    for i, ln in enumerate(raw["lanes"]):
        l = m.lanes.add()
        l.id = i
        l.type = proto.LANE_TYPE_DRIVING
        l.turn = ln["turn"]
        l.max_speed = ln["max_speed"]
        l.length = _length(ln["pts"])
        l.width = LANE_W
        for p in ln["pts"]:
            n = l.center_line.nodes.add()
            n.x, n.y = float(p[0]), float(p[1])
        for s in ln["succ"]:
            c = l.successors.add(); c.id = s; c.type = 1
        for s in ln["pred"]:
            c = l.predecessors.add(); c.id = s; c.type = 2
        l.parent_id = JUNCTION_ID
        l.center_line_type = proto.CENTER_LINE_TYPE_SURFACE_STREET
        junc.lane_ids.append(i)
    for pts, t in raw["lines"]:
        rl = junc.road_lines.add()
        rl.type = t
        for p in pts:
            n = rl.nodes.add(); n.x, n.y = float(p[0]), float(p[1])
    for pts, t in raw["edges"]:
        re = junc.road_edges.add()
        re.type = t
        for p in pts:
            n = re.nodes.add(); n.x, n.y = float(p[0]), float(p[1])
    ag = proto.Agents()
    for a in raw["agents"]:
        p = ag.agents.add()
        p.id = a["id"]
        at = p.attribute
        at.type, at.length, at.width = a["type"], a["length"], a["width"]
        at.max_speed, at.max_acceleration, at.usual_acceleration = 40.0, 5.0, 2.0
        at.max_braking_acceleration, at.usual_braking_acceleration = -5.0, -2.0
        p.home.xy_position.x, p.home.xy_position.y = float(a["xy"][0, 0]), float(a["xy"][0, 1])
        sch = p.schedules.add()
        sch.loop_count = 1
        sch.departure_time = a["k0"] * 0.1
        trip = sch.trips.add()
        trip.mode = proto.TRIP_MODE_DRIVE_ONLY
        trip.end.xy_position.x, trip.end.xy_position.y = float(a["xy"][-1, 0]), float(a["xy"][-1, 1])
        for k in range(len(a["xy"])):
            st = trip.agent_states.add()
            st.position.x, st.position.y = float(a["xy"][k, 0]), float(a["xy"][k, 1])
            st.velocity_x, st.velocity_y = float(a["vel"][k, 0]), float(a["vel"][k, 1])
            st.heading = float(a["heading"][k])
    ag.ads_index = raw["ads_index"]
    return m, ag


def write_scenario(out_dir: str, seed: int, n_agents: int = 128) -> dict:
    raw = make_raw_scenario(seed, n_agents)
    m, ag = to_pb(raw)
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "map.pb"), "wb") as f:
        f.write(m.SerializeToString())
    with open(os.path.join(out_dir, "agents.pb"), "wb") as f:
        f.write(ag.SerializeToString())
    return raw
