"""BASELINE configs[3]: one MOSS-shaped city, 200k vehicles, LaneIDM + spatial-hash collision + off-road every tick.
usage: python profiles/city_bench.py [n_agents] [grid] [ticks]   (prints agent-steps/s of the city engine)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lcsim_b200 import city
from lcsim_b200.city_engine import CityEngine

n_agents = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 70
ticks = int(sys.argv[3]) if len(sys.argv) > 3 else 300
t0 = time.time()
raw = city.make_city(seed=42, grid=grid, n_agents=n_agents, depart_window=10.0, max_hops=6)
st = city.pack_city(raw)
print(f"city: {len(raw['agents'])} vehicles, {st.n_lanes} lanes, {grid}x{grid} junctions; built in {time.time() - t0:.1f} s", flush=True)
ce = CityEngine(st)
ce.step(20)
torch.cuda.synchronize()
ce.reset()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = ce.launches
a.record()
ce.step(ticks)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b)
stats = ce.stats.cpu().numpy()
print(f"{ticks} ticks in {ms:.1f} ms ({1e3 * ms / ticks:.1f} us/tick, {(ce.launches - l0) / ticks:.0f} launches/tick); "
      f"agent-steps {int(stats[0])} -> {stats[0] / (ms * 1e-3):.3e} agent-steps/s; active now {ce.scalar('n_active')}, "
      f"collision sum {int(stats[2])}, off-road sum {int(stats[3])}, arrivals {int(stats[4])}, lc_required {ce.scalar('lc_required')}")
# property check at full size: the spatial-hash collision counts equal brute force on a random sample of agents
xy, h = ce.xy.cpu().numpy(), ce.heading.cpu().numpy()
status, cc = ce.status.cpu().numpy(), ce.coll_count.cpu().numpy()
act = np.nonzero(status == 1)[0]


def _corners(b):  # polygon.py:238-268 corners_from_bboxes
    x, y, length, width, yaw = b[..., 0], b[..., 1], b[..., 2], b[..., 3], b[..., 4]
    cos, sin = np.cos(yaw), np.sin(yaw)
    xc = np.stack([x + length / 2 * cos + width / 2 * sin, x - length / 2 * cos + width / 2 * sin,
                   x - length / 2 * cos - width / 2 * sin, x + length / 2 * cos - width / 2 * sin], axis=-1)
    yc = np.stack([y + length / 2 * sin - width / 2 * cos, y - length / 2 * sin - width / 2 * cos,
                   y - length / 2 * sin + width / 2 * cos, y + length / 2 * sin + width / 2 * cos], axis=-1)
    return np.stack([xc, yc], axis=-1)


def _overlap_a_over_b(first, second):  # polygon.py:206-230 (brute-force checker of this script)
    c, s_ = np.cos(first[..., 4]), np.sin(first[..., 4])
    normals_t = np.stack([np.stack([c, -s_], axis=-1), np.stack([s_, c], axis=-1)], axis=-2)
    pa, pb = np.matmul(_corners(first), normals_t), np.matmul(_corners(second), normals_t)
    return np.all(np.minimum(pa.max(axis=-2), pb.max(axis=-2)) - np.maximum(pa.min(axis=-2), pb.min(axis=-2)) > 0, axis=-1)


rng = np.random.default_rng(0)
sample = rng.choice(act, size=min(300, len(act)), replace=False)
boxes = np.stack([xy[act, 0], xy[act, 1], st.length[act], st.width[act], h[act]], axis=1)
bad = 0
for a_ in sample:
    me = np.asarray([[xy[a_, 0], xy[a_, 1], st.length[a_], st.width[a_], h[a_]]])
    near = np.nonzero((np.abs(boxes[:, 0] - me[0, 0]) < 12) & (np.abs(boxes[:, 1] - me[0, 1]) < 12) & (act != a_))[0]
    n = int(np.logical_and(_overlap_a_over_b(me, boxes[near]), _overlap_a_over_b(boxes[near], me)).sum()) if len(near) else 0
    bad += int(n != cc[a_])
print(f"collision counts vs brute force on {len(sample)} sampled agents: {bad} mismatches")
