"""Experiment (profiles only): where a tick of the rl_env workload (configs[2]) spends its time.

Times, with CUDA events over N back-to-back calls and with the host clock over the same calls (enqueue only), the
pieces of one env tick at S scenarios: step (fused tick + metric item), build_obs (and its parts), rewards, and the
whole tick; a host enqueue time above the device time means the Python / ctypes side is the limit."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", type=int, default=512)
    ap.add_argument("--n", type=int, default=200)
    ap.add_argument("--quick", action="store_true", help="30 warm ticks, then 8 ticks of step / build_obs / rewards and exit (ncu launch list)")
    args = ap.parse_args()
    import numpy as np
    import torch

    from lcsim_b200 import native, workload
    from lcsim_b200.batch import BatchEngine

    torch.cuda.set_device(0)
    S = args.scenarios
    sb, centres = workload.synthetic_batch(100000, S, 128, workers=min(32, os.cpu_count() or 1), with_centres=True)
    be = BatchEngine(sb, reactive=False, with_metrics=True, device=0)
    be.upload_polylines(centres)
    dev = torch.device("cuda", 0)
    act_agent = torch.from_numpy(sb.ads_index.reshape(S, 1).astype(np.int32)).to(dev)
    act_type = torch.full((S, 1), native.ACT_BICYCLE, dtype=torch.int32, device=dev)
    act_data = torch.zeros(S, 1, 5, dtype=torch.float64, device=dev)
    act_data[:, 0, 0] = 0.5
    coef = {"forward": 1.0, "collision": -5.0, "offroad": -5.0, "destination": 10.0}
    be.reset()
    for _ in range(30):
        be.step(act_agent, act_type, act_data)
    o = be.build_obs(k_agents=32, k_polys=64, sort_by_distance=True)
    torch.cuda.synchronize()
    print(json.dumps({"what": "candidates", "n_nbr_agent_mean": float(o["n_nbr_agent"].float().mean()),
                      "n_nbr_poly_mean": float(o["n_nbr_poly"].float().mean()), "n_nbr_poly_max": int(o["n_nbr_poly"].max()),
                      "P": int(be.P), "n_active_mean": float(be.n_active.float().mean())}), flush=True)

    if args.quick:
        ob = rw = None
        for _ in range(8):
            be.step(act_agent, act_type, act_data)
            ob = be.build_obs(k_agents=32, k_polys=64, sort_by_distance=True, out=ob)
            rw = be.rewards(coef=coef, out=rw)
        torch.cuda.synchronize()
        be.close()
        return

    def measure(name, fn, n=args.n):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(json.dumps({"what": name, "device_us_per_call": a.elapsed_time(b) * 1e3 / n, "host_enqueue_us_per_call": (t1 - t0) * 1e6 / n,
                          "wall_us_per_call": (t2 - t0) * 1e6 / n}), flush=True)

    bufs = [None, None]

    def tick():
        be.step(act_agent, act_type, act_data)
        bufs[0] = be.build_obs(k_agents=32, k_polys=64, sort_by_distance=True, out=bufs[0])
        bufs[1] = be.rewards(coef=coef, out=bufs[1])

    ob, rw = [None], [None]

    def obs_reuse():
        ob[0] = be.build_obs(k_agents=32, k_polys=64, sort_by_distance=True, out=ob[0])

    def rew_reuse():
        rw[0] = be.rewards(coef=coef, out=rw[0])

    measure("step", lambda: be.step(act_agent, act_type, act_data), 60)
    be.reset()
    for _ in range(30):
        be.step(act_agent, act_type, act_data)
    measure("build_obs(all)", lambda: be.build_obs(k_agents=32, k_polys=64, sort_by_distance=True))
    measure("build_obs(no history)", lambda: be.build_obs(k_agents=32, k_polys=64, sort_by_distance=True, with_history=False))
    measure("build_obs(no history, no polys)", lambda: be.build_obs(k_agents=32, k_polys=64, sort_by_distance=True, with_history=False, with_polys=False))
    measure("build_obs(unsorted)", lambda: be.build_obs(k_agents=32, k_polys=64, sort_by_distance=False))
    measure("rewards", lambda: be.rewards(coef=coef))
    measure("build_obs(all, out= reused)", obs_reuse)
    for name, kw in (("agents only", dict(with_polys=False, with_history=False)), ("agents + polys", dict(with_history=False)),
                     ("agents + history", dict(with_polys=False)), ("agents + polys unsorted", dict(with_history=False, sort_by_distance=False))):
        args_ = dict(k_agents=32, k_polys=64, sort_by_distance=True)
        args_.update(kw)
        buf = [be.build_obs(**args_)]

        def part(args_=args_, buf=buf):
            buf[0] = be.build_obs(out=buf[0], **args_)

        measure("build_obs(" + name + ", out= reused)", part)
    measure("rewards(out= reused)", rew_reuse)
    be.reset()
    measure("tick = step + build_obs + rewards", tick, 60)
    h_acts = torch.zeros(S, 2, dtype=torch.float64).pin_memory()
    h_out = torch.empty(S * native.ENV_OUT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    be.reset()

    def env_tick():
        be.env_step(h_acts, h_out, coef=coef, sync=False)
        be.build_obs(k_agents=32, k_polys=64, sort_by_distance=True)
        torch.cuda.current_stream(dev).synchronize()

    measure("e2e tick = env_step + build_obs + sync", env_tick, 60)
    be.reset()
    be.set_env_obs(k_agents=32, k_polys=64, sort_by_distance=True)
    measure("e2e tick = env_step (obs registered, sync inside)", lambda: be.env_step(h_acts, h_out, coef=coef, sync=True), 60)
    print(json.dumps({"what": "LCSIM_B200_ENV_ZEROCOPY", "value": os.environ.get("LCSIM_B200_ENV_ZEROCOPY", "default (1)")}), flush=True)
    be.reset()
    d_acts = h_acts.to(dev)
    d_out = [None]

    def dev_tick():
        d_out[0] = be.env_step_device(d_acts, d_out[0], coef=coef)

    measure("device tick = env_step_device (obs registered)", dev_tick, 60)
    be.close()


if __name__ == "__main__":
    main()
