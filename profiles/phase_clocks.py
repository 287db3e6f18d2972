"""Per-phase clock64() breakdown of the tick kernel (instrumented copy, LCSIM_B200_CLOCKS=1).
usage: LCSIM_B200_CLOCKS=1 python profiles/phase_clocks.py [scenarios] [reactive]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lcsim_b200 import workload
from lcsim_b200.batch import BatchEngine

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reactive = len(sys.argv) > 2 and sys.argv[2] == "reactive"
sb = workload.synthetic_batch(0, S, 128)
be = BatchEngine(sb, reactive=reactive, with_metrics=True)
be.lib.lcsim_b200_debug_clocks.restype = C.c_void_p
be.lib.lcsim_b200_debug_clocks.argtypes = [C.c_void_p]
ptr = be.lib.lcsim_b200_debug_clocks(be._h)
assert ptr, "run with LCSIM_B200_CLOCKS=1"
dbg = be.backend.view(ptr, (S, 16), "<i8")
tnames = ["pre", "scan", "post", "depart+publish"] + (["lead search"] if reactive else [])
mnames = ["load snapshot", "collision filter", "edge search", "drain + out"]
NT, NM = len(tnames), len(mnames)
acc_t, acc_m, acc_w = np.zeros(NT), np.zeros(NM), np.zeros(2)
heavy_t, heavy_m = np.zeros(NT), np.zeros(NM)
n = 0
be.rollout(20)
for _ in range(40):
    be.step()
    torch.cuda.synchronize()
    d = dbg.cpu().numpy()
    acc_t += np.diff(d[:, : NT + 1], axis=1).mean(axis=0)
    acc_m += np.diff(d[:, 8 : 8 + NM + 1], axis=1).mean(axis=0)
    acc_w += np.array([d[:, 7].mean(), d[:, 15].mean()])
    tot_t, tot_m = d[:, NT] - d[:, 0], d[:, 8 + NM] - d[:, 8]
    top = np.argsort(tot_t)[-max(1, S // 20):]  # the heaviest 5 % of the tick items: they set the rollout's serial chain
    heavy_t += np.diff(d[top, : NT + 1], axis=1).mean(axis=0)
    heavy_m += np.diff(d[np.argsort(tot_m)[-max(1, S // 20):], 8 : 8 + NM + 1], axis=1).mean(axis=0)
    n += 1
acc_t /= n; acc_m /= n; acc_w /= n; heavy_t /= n; heavy_m /= n
print(f"S={S} reactive={reactive}: mean per-item cycles by phase (thread 0 view, includes barrier waits); one tick per launch")
print(" tick item")
for nm, v, hv in zip(tnames, acc_t, heavy_t):
    print(f"  {nm:18s} {v:9.0f} cyc  {100 * v / acc_t.sum():5.1f}%   heaviest 5%: {hv:9.0f}")
print(f"  total              {acc_t.sum():9.0f} cyc = {acc_t.sum() / 1.965e3:.1f} us @1965 MHz; min/max last tick {tot_t.min()}/{tot_t.max()}; ticket+dependency wait {acc_w[0]:.0f}")
print(" metric item")
for nm, v, hv in zip(mnames, acc_m, heavy_m):
    print(f"  {nm:18s} {v:9.0f} cyc  {100 * v / acc_m.sum():5.1f}%   heaviest 5%: {hv:9.0f}")
print(f"  total              {acc_m.sum():9.0f} cyc = {acc_m.sum() / 1.965e3:.1f} us; min/max last tick {tot_m.min()}/{tot_m.max()}; ticket+dependency wait {acc_w[1]:.0f}")
