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
names = ["pre", "scan", "post", "depart+publish", "lead", "coll filter", "offroad", "coll drain", "out"]
acc = np.zeros(9)
n = 0
be.rollout(20)
for _ in range(40):
    be.step()
    torch.cuda.synchronize()
    d = dbg.cpu().numpy()
    acc += np.diff(d[:, :10], axis=1).mean(axis=0)
    tot = (d[:, 9] - d[:, 0])
    n += 1
acc /= n
print(f"S={S} reactive={reactive}: mean per-CTA cycles by phase (thread 0 view, includes barrier waits)")
for nm, v in zip(names, acc):
    print(f"  {nm:16s} {v:9.0f} cyc  {100 * v / acc.sum():5.1f}%")
print(f"  total            {acc.sum():9.0f} cyc = {acc.sum() / 1.965e3:.1f} us @1965 MHz; CTA total min/max last tick {tot.min()}/{tot.max()}")
