"""Scenario packer: ``map.pb`` / ``agents.pb`` (or a raw synthetic scenario) -> SoA.

Everything the reference computes once in ``Engine.__init__`` and that the hot
path then reads every tick is produced here, in float64, operation for operation
(SURVEY.md H3):

* road-edge polylines resampled at 0.5 m and their per-point direction vectors
  (reference: lcsim/entity/junction/junction.py:69-78 -> utils/polygon.py:90-188,
  and junction.py:374-380 -> polygon.py:78-86),
* per-agent reference trajectories, dimensions, ``dynamic_flag``
  (reference: lcsim/entity/agent/agent.py:133-156),
* the tick at which each agent leaves the waiting list — decided by the engine's
  float64 clock accumulation, not by ``k*0.1`` (reference: engine.py:157,
  agent/manager.py:71-75) — and the waiting-list order
  (manager.py:33-35).

The result (:class:`StaticBatch`) is a set of C-contiguous numpy arrays that is
handed, as plain pointers, to ``lcsim_b200_upload_static`` (include/lcsim_b200.h)
and, identically, to the oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Dict, List, Sequence, Tuple

import numpy as np

from . import proto

NEVER = 2**31 - 1
RESAMPLE_INTERVAL = 0.5  # reference: junction.py:33,63,73


def resample_polyline(pts: np.ndarray, interval: float = RESAMPLE_INTERVAL) -> np.ndarray:
    """Arc-length resampling, reference: polygon.py:90-119 (count) and :142-188 (interp).

    ``n = floor(len/interval)+1`` equally spaced parameters in [0,1]; each is
    located in the normalised cumulative chord table with ``np.digitize`` and
    linearly interpolated inside its segment.  Written with the same numpy
    primitives in the same order so the doubles come out bit-identical.
    """
    pts = np.asarray(pts, dtype=np.float64)
    if pts.shape[0] < 2:
        return pts
    seg = np.linalg.norm(np.diff(pts, axis=0), axis=1)
    n_out = math.floor(float(seg.sum()) / interval) + 1
    u = np.linspace(0, 1, n_out)
    chord = seg / np.sum(seg)
    cum = np.zeros(len(chord) + 1)
    cum[1:] = np.cumsum(chord)
    n = pts.shape[0]
    b = np.digitize(u, bins=cum).astype(int)
    b[np.where((b <= 0) | (u <= 0))] = 1
    b[np.where((b >= n) | (u >= 1))] = n - 1
    frac = np.divide((u - cum[b - 1]), chord[b - 1])
    return pts[b - 1, :] + (pts[b, :] - pts[b - 1, :]) * frac.reshape(-1, 1)


def polyline_dir(pts: np.ndarray) -> np.ndarray:
    """Per-point unit direction of the segment ENDING at the point (first point
    borrows the second's), norm clipped to [1e-6, 1e9]; reference: polygon.py:78-86."""
    prev = np.roll(pts, shift=1, axis=0)
    diff = pts - prev
    if len(diff) > 1:
        diff[0] = diff[1]
    return diff / np.clip(np.linalg.norm(diff, axis=-1)[:, np.newaxis], a_min=1e-6, a_max=1e9)


def depart_tick(departure_time: float, start: float, interval: float, max_ticks: int) -> int:
    """First tick j with ``departure_time <= t_j`` where ``t_{j+1} = t_j + interval``
    is accumulated in float64 exactly like ``Engine.step`` does (engine.py:157)."""
    t = start
    for j in range(max_ticks + 1):
        if departure_time <= t:
            return j
        t += interval
    return NEVER


@dataclass
class ScenarioArrays:
    """One scenario before batching/padding."""

    agent_id: np.ndarray
    agent_type: np.ndarray
    length: np.ndarray
    width: np.ndarray
    departure_time: np.ndarray
    home_xy: np.ndarray
    traj_xy: List[np.ndarray]
    traj_vel: List[np.ndarray]
    traj_h: List[np.ndarray]
    ads_index: int
    edge_xy: np.ndarray  # (E,2) resampled, concatenated in pb order
    edge_dir: np.ndarray  # (E,2)
    edge_poly_start: np.ndarray  # (n_poly+1,) offsets of each polyline in edge_xy
    lanes: list = field(default_factory=list)  # raw centrelines [(pts, lane_id)]
    lines: list = field(default_factory=list)
    edges_raw: list = field(default_factory=list)
    junction_id: int = 0


def _edges_to_arrays(edge_polys: Sequence[np.ndarray]):
    pts, dirs, starts = [], [], [0]
    for p in edge_polys:
        r = resample_polyline(np.asarray(p, dtype=np.float64))
        pts.append(r)
        dirs.append(polyline_dir(r))
        starts.append(starts[-1] + len(r))
    if pts:
        return np.concatenate(pts), np.concatenate(dirs), np.asarray(starts, np.int32)
    return np.zeros((0, 2)), np.zeros((0, 2)), np.asarray(starts, np.int32)


def scenario_from_pb(map_pb, agents_pb) -> ScenarioArrays:
    """WOMD-shaped pb (one junction, xy homes) -> arrays.  reference: engine.py:90-99."""
    if len(map_pb.junctions) != 1:
        raise ValueError("trajectory-agent scenarios need exactly one junction "
                         "(reference: junction/manager.py get_unique_junction)")
    junc = map_pb.junctions[0]
    edge_polys = [np.array([[p.x, p.y] for p in e.nodes]) for e in junc.road_edges]
    edge_xy, edge_dir, starts = _edges_to_arrays(edge_polys)
    ids, types, ls, ws, deps, homes, txy, tvel, th = [], [], [], [], [], [], [], [], []
    for a in agents_pb.agents:
        if a.home.lane_position.ByteSize() > 0 or a.home.xy_position.ByteSize() == 0:
            raise NotImplementedError("lane-position homes (LaneIDM agents) are not packed by this path yet")
        if len(a.schedules) != 1 or len(a.schedules[0].trips) != 1 or a.schedules[0].loop_count != 1:
            raise NotImplementedError("trajectory agents are expected to carry one schedule x one trip, loop_count=1")
        sch = a.schedules[0]
        trip = sch.trips[0]
        if trip.HasField("departure_time"):
            dep = trip.departure_time
        elif sch.HasField("departure_time"):
            dep = sch.departure_time
        elif trip.HasField("wait_time"):
            dep = 0.0 + trip.wait_time
        else:
            dep = 0.0
        if len(trip.agent_states) == 0:
            raise ValueError(f"agent {a.id}: empty agent_states (reference asserts, agent.py:142)")
        ids.append(a.id)
        types.append(a.attribute.type)
        ls.append(a.attribute.length)
        ws.append(a.attribute.width)
        deps.append(dep)
        homes.append([a.home.xy_position.x, a.home.xy_position.y])
        txy.append(np.array([[s.position.x, s.position.y] for s in trip.agent_states]))
        tvel.append(np.array([[s.velocity_x, s.velocity_y] for s in trip.agent_states]))
        th.append(np.array([s.heading for s in trip.agent_states]))
    if len(set(ids)) != len(ids):
        raise ValueError("duplicate agent ids")
    return ScenarioArrays(
        agent_id=np.asarray(ids, np.int32), agent_type=np.asarray(types, np.int32),
        length=np.asarray(ls, np.float64), width=np.asarray(ws, np.float64),
        departure_time=np.asarray(deps, np.float64), home_xy=np.asarray(homes, np.float64).reshape(-1, 2),
        traj_xy=txy, traj_vel=tvel, traj_h=th, ads_index=int(agents_pb.ads_index),
        edge_xy=edge_xy, edge_dir=edge_dir, edge_poly_start=starts,
        lanes=[(np.array([[p.x, p.y] for p in l.center_line.nodes]), l.id) for l in map_pb.lanes],
        lines=[(np.array([[p.x, p.y] for p in l.nodes]), l.type) for l in junc.road_lines],
        edges_raw=[(p, e.type) for p, e in zip(edge_polys, junc.road_edges)],
        junction_id=int(junc.id),
    )


def scenario_from_raw(raw: dict) -> ScenarioArrays:
    """Raw synthetic scenario (synth.make_raw_scenario) -> arrays, same values as via pb."""
    edge_xy, edge_dir, starts = _edges_to_arrays([p for p, _ in raw["edges"]])
    ag = raw["agents"]
    return ScenarioArrays(
        agent_id=np.asarray([a["id"] for a in ag], np.int32),
        agent_type=np.asarray([a["type"] for a in ag], np.int32),
        length=np.asarray([a["length"] for a in ag], np.float64),
        width=np.asarray([a["width"] for a in ag], np.float64),
        departure_time=np.asarray([a["k0"] * 0.1 for a in ag], np.float64),
        home_xy=np.asarray([a["xy"][0] for a in ag], np.float64).reshape(-1, 2),
        traj_xy=[a["xy"] for a in ag], traj_vel=[a["vel"] for a in ag], traj_h=[a["heading"] for a in ag],
        ads_index=int(raw["ads_index"]), edge_xy=edge_xy, edge_dir=edge_dir, edge_poly_start=starts,
        lanes=[(ln["pts"], i) for i, ln in enumerate(raw["lanes"])],
        lines=list(raw["lines"]), edges_raw=list(raw["edges"]), junction_id=300000000,
    )


def _chunk_centres(pts: np.ndarray, chunk: int) -> np.ndarray:
    """(x, y, heading) of every `chunk`-point piece of one resampled polyline: the float32 mean of
    [x, y, z, dir_x, dir_y, dir_z] (reference: polygon.py:41-50) and atan2(dir_y, dir_x) in float32
    (motion_diff/modules/agent_encoder.py:125-134)."""
    if len(pts) == 0:
        return np.zeros((0, 3))
    with_z = np.concatenate([pts, np.zeros((len(pts), 1))], axis=1)
    rows = np.concatenate([with_z, polyline_dir(with_z)], axis=1)
    out = []
    for i in range(0, len(rows), chunk):
        c = rows[i:i + chunk].mean(axis=0).astype(np.float32)
        out.append([float(c[0]), float(c[1]), float(np.arctan2(c[4], c[3]))])
    return np.asarray(out)


def polyline_centres(lanes) -> np.ndarray:
    """(P,3) polyline centres (x, y, heading) of a WOMD-shaped scenario, in the order
    Junction.init_roadgraph_data concatenates them (junction/junction.py:194-205,231-236): only the
    junction's LANES contribute (0.5 m-resampled centrelines cut into 10-point pieces, lane/lane.py:59-63,
    108-124); the junction's road lines / edges go to `roadgraph_points`, not to the polyline set.
    `lanes` = ScenarioArrays.lanes = [(centre-line nodes (n,2), lane id)] in pb order."""
    parts = [_chunk_centres(resample_polyline(np.asarray(p, np.float64)), 10) for p, _ in lanes]
    parts = [p for p in parts if len(p)]
    return np.concatenate(parts) if parts else np.zeros((0, 3))


def centreline_points(lanes, lines=()):
    """The points the reference's lane projections search: Junction.roadgraph_points with type 1|2
    (motion_diff/modules/guidance.py:110-111, metrics/roadgraph.py:139-142), in the order
    Junction.init_roadgraph_data concatenates them (junction/junction.py:194-205,240-253): per lane the 0.5 m-resampled
    centre-line (lane/lane.py:108-124; type CENTER_LINE_TYPE_SURFACE_STREET = 2, polyline id = lane ordinal), then —
    a quirk of the reference's type test — every junction ROAD LINE whose raw pb type is 1 or 2 (junction.py:122-132
    stores road-line types without a prefix), polyline id = number of lanes + line ordinal.  float32 xyz and float32
    get_polyline_dir.  `lanes` = ScenarioArrays.lanes, `lines` = ScenarioArrays.lines.  Returns
    (xy_theta (N,3) float32 with theta = atan2(dir_y, dir_x) in float32 like roadgraph.py:149, ids (N,) int32,
    xy_dir (N,4) float32 for parity checks)."""
    xs, ds, ids = [], [], []
    polys = [(p, i) for i, (p, _) in enumerate(lanes)]
    polys += [(p, len(lanes) + j) for j, (p, t) in enumerate(lines) if int(t) in (1, 2)]
    for p, pid in polys:
        c = resample_polyline(np.asarray(p, np.float64))
        if len(c) == 0:
            continue
        with_z = np.concatenate([c, np.zeros((len(c), 1))], axis=1)
        xs.append(c.astype(np.float32))
        ds.append(polyline_dir(with_z)[:, :2].astype(np.float32))
        ids.append(np.full(len(c), pid, np.int32))
    if not xs:
        return np.zeros((0, 3), np.float32), np.zeros(0, np.int32), np.zeros((0, 4), np.float32)
    xy, d = np.concatenate(xs), np.concatenate(ds)
    theta = np.arctan2(d[:, 1], d[:, 0]).astype(np.float32)
    return np.concatenate([xy, theta[:, None]], axis=1).astype(np.float32), np.concatenate(ids), np.concatenate([xy, d], axis=1)


@dataclass
class StaticBatch:
    """Padded SoA for S scenarios; layout documented in DESIGN.md §3 and include/lcsim_b200.h."""

    S: int
    A: int
    T: int
    E: int
    start: float
    interval: float
    total_steps: int
    n_agents: np.ndarray  # (S,) i32
    ads_index: np.ndarray  # (S,) i32
    agent_id: np.ndarray  # (S,A) i32
    agent_type: np.ndarray  # (S,A) i32
    length: np.ndarray  # (S,A) f64
    width: np.ndarray  # (S,A) f64
    dynamic: np.ndarray  # (S,A) u8
    traj_len: np.ndarray  # (S,A) i32
    depart_tick: np.ndarray  # (S,A) i32
    wait_order: np.ndarray  # (S,A) i32
    home_xy: np.ndarray  # (S,A,2) f64
    traj_xy: np.ndarray  # (S,A,T,2) f64
    traj_vel: np.ndarray  # (S,A,T,2) f64
    traj_h: np.ndarray  # (S,A,T) f64
    n_edge: np.ndarray  # (S,) i32
    edge_xy: np.ndarray  # (S,E,2) f64
    edge_dir: np.ndarray  # (S,E,2) f64
    origin: np.ndarray  # (S,2) f64
    edge_poly_id: Optional[np.ndarray] = None  # (S,E) i32 ordinal of the road-edge polyline of every point (roadgraph ids)

    def scenario_slice(self, s0: int, s1: int) -> "StaticBatch":
        kw = {k: (v[s0:s1] if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items()}
        kw["S"] = s1 - s0
        return StaticBatch(**{k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in kw.items()})


def pack(scenarios: Sequence[ScenarioArrays], start: float = 0.0, interval: float = 0.1,
         total_steps: int = 91, A: int | None = None, T: int | None = None, E: int | None = None) -> StaticBatch:
    S = len(scenarios)
    A = A or max(len(s.agent_id) for s in scenarios)
    T = T or max(max((len(x) for x in s.traj_xy), default=1) for s in scenarios)
    E = E or max(1, max(len(s.edge_xy) for s in scenarios))
    # pad E to a whole number of 32-point blocks (the edge tile is scanned block-wise)
    E = (E + 31) // 32 * 32
    b = StaticBatch(
        S=S, A=A, T=T, E=E, start=float(start), interval=float(interval), total_steps=int(total_steps),
        n_agents=np.zeros(S, np.int32), ads_index=np.zeros(S, np.int32),
        agent_id=np.full((S, A), -1, np.int32), agent_type=np.zeros((S, A), np.int32),
        length=np.zeros((S, A)), width=np.zeros((S, A)), dynamic=np.zeros((S, A), np.uint8),
        traj_len=np.zeros((S, A), np.int32), depart_tick=np.full((S, A), NEVER, np.int32),
        wait_order=np.zeros((S, A), np.int32), home_xy=np.zeros((S, A, 2)),
        traj_xy=np.zeros((S, A, T, 2)), traj_vel=np.zeros((S, A, T, 2)), traj_h=np.zeros((S, A, T)),
        n_edge=np.zeros(S, np.int32), edge_xy=np.zeros((S, E, 2)), edge_dir=np.zeros((S, E, 2)),
        origin=np.zeros((S, 2)), edge_poly_id=np.zeros((S, E), np.int32),
    )
    max_ticks = max(4 * total_steps, 1024)
    for i, sc in enumerate(scenarios):
        n = len(sc.agent_id)
        if n > A:
            raise ValueError(f"scenario {i}: {n} agents > A={A}")
        if len(sc.edge_xy) > E:
            raise ValueError(f"scenario {i}: {len(sc.edge_xy)} edge points > E={E}")
        b.n_agents[i] = n
        b.ads_index[i] = sc.ads_index
        b.agent_id[i, :n] = sc.agent_id
        b.agent_type[i, :n] = sc.agent_type
        b.length[i, :n] = sc.length
        b.width[i, :n] = sc.width
        b.home_xy[i, :n] = sc.home_xy
        for a in range(n):
            xy = np.asarray(sc.traj_xy[a], np.float64)
            t = len(xy)
            if t > T:
                raise ValueError(f"scenario {i} agent {a}: {t} states > T={T}")
            b.traj_len[i, a] = t
            b.traj_xy[i, a, :t] = xy
            b.traj_vel[i, a, :t] = sc.traj_vel[a]
            b.traj_h[i, a, :t] = sc.traj_h[a]
            # reference: agent.py:153-156 (bbox diagonal of the whole trajectory > 1 m and a vehicle)
            diag = np.linalg.norm(np.abs(np.max(xy, axis=0) - np.min(xy, axis=0)))
            b.dynamic[i, a] = 1 if (diag > 1.0 and sc.agent_type[a] == proto.AGENT_TYPE_VEHICLE) else 0
            b.depart_tick[i, a] = depart_tick(float(sc.departure_time[a]), start, interval, max_ticks)
        # reference: manager.py:33-35 — sorted() is stable, so ties keep pb order
        order = sorted(range(n), key=lambda k: sc.departure_time[k])
        b.wait_order[i, :n] = order
        ne = len(sc.edge_xy)
        b.n_edge[i] = ne
        b.edge_xy[i, :ne] = sc.edge_xy
        b.edge_dir[i, :ne] = sc.edge_dir
        st = np.asarray(sc.edge_poly_start, np.int64)
        if len(st) > 1 and int(st[-1]) == ne:  # (tests build scenarios whose edge arrays were cut by hand)
            b.edge_poly_id[i, :ne] = np.repeat(np.arange(len(st) - 1, dtype=np.int32), np.diff(st))
        # origin of the fp32 filter frame: integer metres near the middle of everything
        allp = [sc.edge_xy] if ne else []
        allp += [np.asarray(x, np.float64) for x in sc.traj_xy]
        allp = np.concatenate(allp) if allp else np.zeros((1, 2))
        b.origin[i] = np.round(0.5 * (allp.min(axis=0) + allp.max(axis=0)))
    return b


def pack_raw(raws: Sequence[dict], **kw) -> StaticBatch:
    return pack([scenario_from_raw(r) for r in raws], **kw)


def scenarios_from_dirs(dirs: Sequence[str]) -> List[ScenarioArrays]:
    import os

    return [scenario_from_pb(proto.load_map(os.path.join(d, "map.pb")), proto.load_agents(os.path.join(d, "agents.pb")))
            for d in dirs]


def pack_dirs(dirs: Sequence[str], **kw) -> StaticBatch:
    return pack(scenarios_from_dirs(dirs), **kw)
