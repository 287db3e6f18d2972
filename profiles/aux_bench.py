"""Timing of the kernels around the tick on synthetic WOMD-shaped scenarios, CUDA
events on the launching stream after warm-up, L2 flushed between repetitions.
usage: python profiles/aux_bench.py [scenarios=512]      (one JSON line per kernel)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lcsim_b200 import workload
from lcsim_b200.batch import BatchEngine

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for k in range(reps):
        flush.fill_(k)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def report(name, ms, bytes_alg, what):
    print(json.dumps({"kernel": name, "scenarios": S, "ms": round(ms, 4), "algorithmic_GB": round(bytes_alg / 1e9, 4),
                      "GBps": round(bytes_alg / (ms * 1e-3) / 1e9, 1), "what": what}), flush=True)


sb, centres, lanes = workload.synthetic_batch(0, S, 128, with_centres="lanes")
N_rg = max(len(x[0]) for x in lanes)
if len(sys.argv) > 2:  # child process: the plain-load lane kernel (the switch is read once per process)
    os.environ["LCSIM_B200_LANE_BULK"] = "0"
    be = BatchEngine(sb, reactive=False, with_metrics=True, device=0)
    be.upload_centrelines(lanes)
    be.rollout(30)
    ms = timed(lambda: be.project_lanes(top_k=64))
    report("lcs_lane_kernel<plain> (LCSIM_B200_LANE_BULK=0: CTA loads / stores)", ms, S * 16.0 * N_rg, "16 B per centreline point per scenario")
    be.close()
    sys.exit(0)
os.environ.pop("LCSIM_B200_LANE_BULK", None)
be = BatchEngine(sb, reactive=False, with_metrics=True, device=0)
be.upload_polylines(centres)
be.upload_centrelines(lanes)
be.rollout(30)
n_act = int(be.n_active.sum().item())
P = int(be.P)
ms = timed(lambda: be.build_obs(k_agents=32, k_polys=64, sort_by_distance=True))
report("lcs_obs_kernel (ego: top-32 agents, top-64 polylines, history stack)", ms,
       S * (12.0 * P + 16.0 * 96) + n_act * 11 * (40 + 24), "12 B per polyline centre + 16 B per neighbour + history in/out")
ms = timed(lambda: be.build_encoder_inputs())
report("lce_encoder_kernel (all agents: x_a, r_t, a2a CSR, pl2a CSR)", ms,
       S * 12.0 * P + n_act * 11 * (40 + 24 + 16), "12 B per polyline centre + history in + features out (edges extra)")
ms = timed(lambda: be.project_lanes(top_k=64))
report("lcs_lane_kernel<bulk> (cp.async.bulk + mbarrier tiles)", ms, S * 16.0 * N_rg, "16 B per centreline point per scenario")
g = torch.Generator(device=dev)
g.manual_seed(0)
T = 80
xy = (be.pose64[:, :, None, :2] + torch.randn(S, be.A, T, 2, generator=g, device=dev, dtype=torch.float64).cumsum(2)).float()
ms = timed(lambda: be.guidance_cost("repeller", xy))
report("lcp_repeller_kernel + finalize (128 agents x 80 steps, loss + gradient)", ms, 2 * xy.numel() * 4, "plans in + gradient out")
ms = timed(lambda: be.guidance_cost("no_offroad", xy))
report("lcp_offroad_kernel + finalize (NoOffRoad loss + gradient)", ms, 2 * xy.numel() * 4 + S * 16.0 * sb.E, "plans + gradient + 16 B per edge point")
traj5 = torch.cat([xy, torch.full((S, be.A, T, 2), 4.5, device=dev), torch.zeros(S, be.A, T, 1, device=dev)], dim=-1).contiguous()
traj5[..., 3] = 2.0
ms = timed(lambda: be.plan_metrics(traj5))
report("lcp_overlap_kernel + lcp_offroad_kernel (overlap rate, off-road rate)", ms, traj5.numel() * 4 * 2, "plans read twice")
be.close()
import subprocess

subprocess.run([sys.executable, os.path.abspath(__file__), str(S), "plain-lane"], check=False)
