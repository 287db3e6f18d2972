"""Experiment harness (profiles only): one synthetic batch, several engine variants, device-resident rollout throughput.

    python profiles/exp_variants.py [--scenarios 1024] [--rollouts 20] VARIANT ...

VARIANT = name:mode:ENV=V,ENV=V   (mode: replay | reactive); the environment variables are the engine's experiment
knobs read at create() (LCSIM_B200_ORDER, LCSIM_B200_SNAP_DEPTH, ...).  LCSIM_B200_LIB selects another build of the
library and therefore needs its own process.  Prints one JSON line per variant; the end state of every variant's
last rollout (poses, statuses, collision counts, nearest edges) is hashed so that variants can be compared bit for bit.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", type=int, default=1024)
    ap.add_argument("--agents", type=int, default=128)
    ap.add_argument("--rollouts", type=int, default=20)
    ap.add_argument("--ticks", type=int, default=91)
    ap.add_argument("variants", nargs="+")
    args = ap.parse_args()
    import torch

    from lcsim_b200 import workload
    from lcsim_b200.batch import BatchEngine

    torch.cuda.set_device(0)
    t0 = time.time()
    sb = workload.synthetic_batch(0, args.scenarios, args.agents, workers=min(32, os.cpu_count() or 1))
    gen_s = time.time() - t0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for v in args.variants:
        name, mode, envs = (v.split(":") + ["", ""])[:3]
        env = dict(kv.split("=") for kv in envs.split(",") if kv)
        saved = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        t0 = time.time()
        be = BatchEngine(sb, reactive=mode == "reactive", with_metrics=True, device=0)
        up_s = time.time() - t0
        for k, val in saved.items():
            if val is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = val
        for _ in range(3):
            be.reset()
            be.rollout(args.ticks)
        torch.cuda.synchronize()
        ms = []
        for r in range(args.rollouts):
            flush.fill_(r & 0xFF)
            be.reset()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            be.rollout(args.ticks)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        steps = int(be.backend.to_numpy(be.episode_stats())[0])  # statistics restart at every reset: the last rollout's
        h = hashlib.sha256()
        for nm in ("pose64", "status", "coll_count", "nearest_edge", "offroad", "cursor"):
            h.update(be.backend.to_numpy(getattr(be, nm)).tobytes())
        ms.sort()
        med = ms[len(ms) // 2]
        print(json.dumps({"variant": name, "mode": mode, "env": env, "agent_steps_per_rollout": steps, "ms_median": med,
                          "ms_min": ms[0], "ms_max": ms[-1], "agent_steps_per_s": steps / (med * 1e-3),
                          "state_sha": h.hexdigest()[:16], "gen_s": round(gen_s, 2), "upload_s": round(up_s, 2)}), flush=True)
        be.close()


if __name__ == "__main__":
    main()
