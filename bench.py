#!/usr/bin/env python
"""bench.py — agent-steps/s of the per-tick agent-stepping path on synthetic WOMD-shaped scenarios.

A "step" is one 91-tick rollout (reset launch + one fused launch that runs the 91 ticks) of this rank's batch of
independent scenarios (BASELINE.json configs[1]: 1024 scenarios x 128 agents on one B200; log-replay
control with the collision / off-road metrics fused into every tick).  agent-steps = sum over ticks of the
ACTIVE agents updated.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework
  python bench.py --impl reference ...                           # the CPU reference arm (see below)

Reference arm: the reference is pure Python (no compiled source to build, and /root/reference does not
travel to the GPU box), so `--impl reference` times oracle/lcsim_oracle.c — the C restatement of the
reference's algorithm that tests/ pin to golden vectors of the real engine — on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TICKS = 91
B_ALG = 1356  # algorithmic bytes per agent-step at A=128, T=91, E=4096
METRIC = "agent-steps/sec"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_rollouts(sb, threads: int, n_rollouts: int = 1):
    """The reference arm / cpu_baseline: oracle (C restatement of the reference) on `threads` host threads.
    Returns (agent_steps, seconds)."""
    from oracle import oracle as orc

    o = orc.Oracle(sb, reactive=False)
    steps, t = 0, 0.0
    for _ in range(n_rollouts):
        o.reset()
        t0 = time.perf_counter()
        steps += o.rollout_mt(N_TICKS, True, threads)
        t += time.perf_counter() - t0
    return steps, t


def run_reference(args):
    """bench.py --impl reference: rank 0 only; each step = one rollout of a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lcsim_b200 import workload

    cores = os.cpu_count() or 1
    n_sample = args.cpu_scenarios
    sb = workload.synthetic_batch(0, n_sample, args.agents)
    for _ in range(args.warmup):
        cpu_rollouts(sb.scenario_slice(0, min(n_sample, cores)), cores)
    steps, secs = cpu_rollouts(sb, cores, args.steps)
    val = steps / secs
    sample = f"{n_sample} of the {args.scenarios} scenarios per step, 91 ticks with collision/off-road metrics"
    print_json({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(args, "host memory"),
        "cpu_baseline": {"value": val, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is pure Python and cannot travel; timed arm is oracle/lcsim_oracle.c (pinned to the real "
                "engine's golden vectors) on all host cores",
    })


def _config(args, l2):
    return {"workload": f"{args.scenarios} synthetic WOMD-shaped scenarios per GPU x {args.agents} agents, T<=91, "
                        f"~3.9k road-edge points, 91-tick log-replay rollout with fused collision/off-road metrics "
                        f"(BASELINE configs[1])",
            "scenarios_per_gpu": args.scenarios, "agents": args.agents, "ticks": N_TICKS,
            "mode": "log-replay" if args.mode == "replay" else "reactive (TrajIDM + bicycle)",
            "l2": l2,
            "exact_shortcuts": "replay records (cursor of a replayed waypoint from a canonical-index table instead of the "
                               "trajectory scan; LCSIM_B200_RECORDS=0 disables), edge search skipped for agents whose position "
                               "did not change, per-cell candidate lists for the nearest road-edge point "
                               "(LCSIM_B200_CAND_CELL=0 disables); results bit-identical either way (tests/test_parity.py)"}


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from lcsim_b200 import native, shard, workload
    from lcsim_b200.batch import BatchEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    S = args.scenarios
    sb = workload.synthetic_batch(rank * S, S, args.agents, workers=max(1, min(32, (os.cpu_count() or 1) // world)))
    be = BatchEngine(sb, reactive=(args.mode == "reactive"), with_metrics=not args.no_metrics, device=local,
                     grid_cell=args.grid_cell)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ------------------------------------------------------------- device-resident arm ("value")
    kev = []  # CUDA events around the rollout kernel alone (the roofline's launch duration)

    def rollout(timed=False):
        be.reset()
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        be.rollout(N_TICKS)
        if timed:
            b.record()
            kev.append((a, b))

    for _ in range(args.warmup):
        rollout()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = be.launches
    t_wall0 = time.time()
    agent_steps = 0
    stats_acc = torch.zeros(native.N_STATS, dtype=torch.int64, device=dev)
    for k in range(args.steps):
        flush.fill_(k)  # L2 flushed between timed rollouts (not timed)
        ev[k][0].record()
        rollout(timed=True)
        ev[k][1].record()
        stats_acc += be.episode_stats()
    barrier()
    t_wall1 = time.time()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = be.launches - launches0
    total = shard.reduce_episode_stats(stats_acc)  # the one collective of the path: NCCL all-reduce over NVLink
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    agent_steps_all = int(total[0].item())
    agent_steps_rank = int(stats_acc[0].item())
    value = agent_steps_all / (ms_max * 1e-3)

    e2e_val, h2d, d2h = None, 0, 0
    roll_val, roll_h2d, roll_d2h = None, 0, 0
    if not args.kernel_only:
        # ------------------------------------------------------------- end-to-end arm ("e2e"): host buffers, one call
        # The reference-facing call for THIS workload (configs[1]: `engine.reset(); for _ in range(91): engine.step()`
        # on every scenario, nothing fed in between): lcsim_b200_rollout_host.  Per step (= rollout): H2D of the reset
        # mask from pinned memory, reset + 91 ticks, D2H of the per-tick log, float32 poses, status, collision counts,
        # nearest edge, off-road flags and the statistics into pinned memory, stream sync.
        Ap = be.A_stride
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()  # noqa: E731
        r_out = dict(tick_log=pin((N_TICKS, S, 4), torch.int32), pose32=pin((S, Ap, 4), torch.float32),
                     status=pin((S, Ap), torch.int32), coll_count=pin((S, Ap), torch.int32),
                     nearest_edge=pin((S, Ap), torch.int32), offroad=pin((S, Ap), torch.uint8), stats=pin((8,), torch.int64))
        r_mask = torch.ones(S, dtype=torch.uint8).pin_memory()
        roll_h2d = r_mask.numel()
        roll_d2h = sum(v.numel() * v.element_size() for v in r_out.values())
        for _ in range(min(args.warmup, 3)):
            be.rollout_host(N_TICKS, r_out, reset_mask=r_mask)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        roll_steps = 0
        r0.record()
        for k in range(args.steps):
            be.rollout_host(N_TICKS, r_out, reset_mask=r_mask)  # returns after the stream sync: host buffers are final
            roll_steps += int(r_out["stats"][0])
        r1.record()
        barrier()
        t = torch.tensor([r0.elapsed_time(r1), float(roll_steps)], dtype=torch.float64, device=dev)
        if world > 1:
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t[0] = tm[0]
        roll_val = float(t[1]) / (float(t[0]) * 1e-3)

        # ------------------------------------------------------------- end-to-end arm ("e2e"): host buffers
        # The reference-facing call: lcsim_b200_env_step == BicycleSingleEnv.step for all S scenarios.  Every tick the
        # normalised ADS action of every scenario goes host -> device from pinned memory and reward / terms / route /
        # terminated / truncated come back device -> host; the host reads them (stream sync) before the next tick,
        # like an RL loop does.  Every rollout ends with the D2H of the state a host caller reads.
        rng = np.random.default_rng(1)
        acts = torch.from_numpy(np.stack([rng.uniform(-0.3, 0.3, (N_TICKS, S)), rng.uniform(-0.2, 0.2, (N_TICKS, S))],
                                         axis=-1)).contiguous().pin_memory()
        h_out = torch.empty(S * native.ENV_OUT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        out_np = h_out.numpy().view(native.ENV_OUT_DTYPE)
        coef = {"forward": 2, "collision": 5, "smooth": 1, "destination": 1}
        h_pose = torch.empty(S, be.A, 4, dtype=torch.float32).pin_memory()
        h_int = torch.empty(3, S, be.A, dtype=torch.int32).pin_memory()
        h_stats = torch.empty(native.N_STATS, dtype=torch.int64).pin_memory()
        h2d = d2h = 0
        ret = 0.0

        def e2e_rollout(count=False):
            nonlocal h2d, d2h, ret
            be.reset()
            for k in range(N_TICKS):
                be.env_step(acts[k], h_out, coef=coef, sync=True)
                ret += float(out_np["reward"][0])  # the host consumes the result
                if count:
                    h2d += acts[k].numel() * 8
                    d2h += h_out.numel()
            h_pose.copy_(be.pose32, non_blocking=True)
            h_int[0].copy_(be.status, non_blocking=True)
            h_int[1].copy_(be.coll_count, non_blocking=True)
            h_int[2].copy_(be.nearest_edge, non_blocking=True)
            h_stats.copy_(be.episode_stats(), non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            if count:
                d2h += h_pose.numel() * 4 + h_int.numel() * 4 + h_stats.numel() * 8
            return int(h_stats[0])

        for _ in range(min(args.warmup, 3)):
            e2e_rollout()
        barrier()
        e2e_steps = 0
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.steps):
            e2e_steps += e2e_rollout(count=(k == 0))
        e1.record()
        barrier()
        e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3 * 0)  # device clock around host-driven loop
        t = torch.tensor([e2e_ms, float(e2e_steps)], dtype=torch.float64, device=dev)
        if world > 1:
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            e2e_ms, e2e_steps = float(tm[0]), int(t[1])
        e2e_val = e2e_steps / (e2e_ms * 1e-3)

    # clocks / throttle reasons over everything timed on the GPU (the device-resident leg and both end-to-end legs)
    clocks = sampler.stop(t_wall0, time.time() if not args.kernel_only else t_wall1)
    if rank == 0:
        peak, which = _peaks()
        # dominant kernel: the stepping launches of a rollout (one fused lcs_rollout_kernel launch of 91 ticks, or 91
        # lcs_tick_kernel launches with LCSIM_B200_FUSED=0), timed by their own CUDA events inside the timed region
        n_step_launches = max(launches - args.steps, 1)  # every rollout also has one reset launch
        per_launch_steps = agent_steps_rank / n_step_launches
        launch_ms = sum(a.elapsed_time(b) for a, b in kev) / n_step_launches
        achieved = per_launch_steps * B_ALG / (launch_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "tick_kernel_traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_scenarios
        cpu_steps, cpu_secs = (0, 1.0) if args.kernel_only else cpu_rollouts(sb.scenario_slice(0, n_cpu), cores)
        out = {
            "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": _config(args, "inputs 5.6 GB/GPU > 126 MB L2; L2 flushed between timed rollouts"),
            "clocks": clocks,
            "e2e": {"value": roll_val, "unit": "agent-steps/s", "h2d_bytes_per_step": roll_h2d, "d2h_bytes_per_step": roll_d2h,
                    "what": "lcsim_b200_rollout_host, one call per rollout with pinned host buffers: H2D reset mask, reset + 91 "
                            "ticks, D2H of the per-tick log [91][S][4], pose32, status, collision counts, nearest edge, "
                            "off-road flags, statistics; stream sync before the host reads them"},
            "e2e_env_step": {"value": e2e_val, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                             "what": "RL-style loop (configs[2] call pattern on the same batch): lcsim_b200_env_step per tick "
                                     "(pinned H2D of every scenario's ADS action, tick, reward kernel, D2H of reward/terms/"
                                     "route/flags, stream sync every tick) + per rollout D2H of pose32/status/collision "
                                     "counts/nearest-edge/statistics"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": which,
                         "kernel": "lcs_rollout_kernel" if n_step_launches == args.steps else "lcs_tick_kernel",
                         "bytes_per_agent_step": B_ALG, "agent_steps_per_launch": per_launch_steps,
                         "launch_ms": launch_ms},
            "cpu_baseline": {"value": cpu_steps / cpu_secs, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                             "sample": f"first {n_cpu} scenarios of rank 0's batch, one 91-tick rollout with metrics, "
                                       f"oracle/lcsim_oracle.c on {cores} threads"},
            "agent_steps_per_rollout": agent_steps_all // max(args.steps, 1),
            "episode": shard.rates(total),
        }
        print_json(out)
    be.close()
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """Libraries (NCCL's version banner, torchrun warnings) write to fd 1; the contract is ONE JSON line on stdout.
    Everything else goes to stderr: fd 1 is pointed at stderr and the JSON line is written to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(saved, "w")
    global print_json
    print_json = lambda obj: (out.write(json.dumps(obj) + "\n"), out.flush())  # noqa: E731


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scenarios", type=int, default=1024, help="scenarios per GPU")
    ap.add_argument("--agents", type=int, default=128)
    ap.add_argument("--mode", default="replay", choices=["replay", "reactive"], help="control mode of the timed rollout")
    ap.add_argument("--grid-cell", type=float, default=0.0, help="experiment: road-edge grid cell (m); 0 = library default")
    ap.add_argument("--no-metrics", action="store_true", help="experiment: skip the fused collision/off-road phases")
    ap.add_argument("--kernel-only", action="store_true", help="skip the e2e and CPU legs (profiling runs)")
    ap.add_argument("--cpu-scenarios", type=int, default=0, help="bounded CPU sample (scenarios per rollout); "
                    "0 = max(64, 2 x host cores)")
    args = ap.parse_args()
    if args.cpu_scenarios <= 0:
        args.cpu_scenarios = min(args.scenarios, max(64, 2 * (os.cpu_count() or 1)))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
