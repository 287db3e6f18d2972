#!/usr/bin/env python
"""bench.py — agent-steps/s of the per-tick agent-stepping path on synthetic WOMD-shaped scenarios.

Headline (BASELINE.json configs[1]): 1024 scenarios x 128 agents per GPU, 91-tick log-replay rollouts with the
collision / off-road metrics of every tick.  A "step" is `--rollouts-per-step` (default 50) consecutive rollouts
(reset + 91 ticks, ONE kernel launch each) of this rank's batch, so that the `--steps 20` the driver asks for give
a timed region of seconds.  agent-steps = sum over ticks of the ACTIVE agents updated.

The same JSON line carries one sub-record per other BASELINE config / control mode, each with its own value,
roofline, e2e and clocks:  `reactive` (TrajIDM car-following, same batch), `rl_env` (configs[2]: 512 envs per GPU,
env_step + observation build + rewards every tick), `replan` (configs[4]: 512 scenarios, float32 re-plans fed in at
the reference's plan schedule), `city` (configs[3]: one MOSS-shaped city with 200k vehicles, LaneIDM).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework, all workloads
  python bench.py --workload reactive ...                        # one workload as the headline (experiments)
  python bench.py --impl reference ...                           # the CPU reference arm (see below)

Reference arm: the reference is pure Python (no compiled source to build, and /root/reference does not
travel to the GPU box), so `--impl reference` times oracle/lcsim_oracle.c — the C restatement of the
reference's algorithm that tests/ pin to golden vectors of the real engine — on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TICKS = 91
METRIC = "agent-steps/sec"
WORKLOADS = ("replay", "reactive", "rl_env", "replan", "city")
PLAN_INIT, PLAN_INTERVAL, PLAN_T = 10, 80, 80  # lcsim/config/example.yml: plan_init_step, plan_interval; 80-step plans
WAYMO_PPO_REWARD = {"forward": 0.1, "collision": 10, "offroad": 5, "smooth": 2, "destination": 1}  # waymo_ppo.yml:36-41
B_CITY = 420  # algorithmic bytes per agent-step of the city tick


def b_alg(T=91, E=4096, A=128, fp64_master=False, obs_bytes_per_agent_step=0.0):
    """B_alg = 116 + 8 T + 16 E / A (+32 with an fp64 master pose, + observation terms)."""
    return 116 + 8 * T + 16 * E / A + (32 if fp64_master else 0) + obs_bytes_per_agent_step


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

    def window(self, t0, t1):
        """clocks over [t0, t1] (wall clock)"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows[-3:]]
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nm, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()


def cpu_rollouts(sb, threads: int, n_rollouts: int = 1, reactive: bool = False):
    """The reference arm / cpu_baseline: oracle (C restatement of the reference) on `threads` host threads.
    Returns (agent_steps, seconds)."""
    from oracle import oracle as orc

    o = orc.Oracle(sb, reactive=reactive)
    steps, t = 0, 0.0
    for _ in range(n_rollouts):
        o.reset()
        t0 = time.perf_counter()
        steps += o.rollout_mt(N_TICKS, True, threads)
        t += time.perf_counter() - t0
    return steps, t


def _workload_text(args):
    return (f"{args.scenarios} synthetic WOMD-shaped scenarios per GPU x {args.agents} agents, T<=91, ~3.9k road-edge "
            f"points, 91-tick log-replay rollout with collision/off-road metrics every tick (BASELINE configs[1]); "
            f"step = {args.rollouts_per_step} consecutive rollouts (reset + one launch of 91 ticks each)")


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
    sample = f"{n_sample} of the {args.scenarios} scenarios per step, one 91-tick rollout with collision/off-road metrics"
    print_json({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": _workload_text(args), "scenarios_per_gpu": args.scenarios, "agents": args.agents,
                   "ticks": N_TICKS, "mode": "log-replay", "l2": "host memory"},
        "cpu_baseline": {"value": val, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is pure Python and cannot travel; timed arm is oracle/lcsim_oracle.c (pinned to the real "
                "engine's golden vectors) on all host cores",
    })


# ----------------------------------------------------------------------------------------------------------------------


class Ctx:
    """rank / device / collectives / L2 flush shared by the workloads"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.sampler = ClockSampler(self.local)
        self.sampler.start()

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_sum(self, ms: float, count: float):
        """(max over ranks of ms, sum over ranks of count)"""
        t = self.torch.tensor([ms, count], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            tm = t.clone()
            self.dist.all_reduce(tm, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            t[0] = tm[0]
        return float(t[0]), float(t[1])

    def timed(self, step_fn, steps, warmup, flush=True):
        """W warm-up steps, then K steps, each bracketed by CUDA events on the current stream, L2 flushed between the
        steps (untimed); barrier + synchronize on both sides.  Returns (sum of step times in ms, wall window)."""
        torch = self.torch
        for _ in range(warmup):
            step_fn()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        w0 = time.time()
        for k in range(steps):
            if flush:
                self.flush.fill_(k & 0xFF)
            ev[k][0].record()
            step_fn()
            ev[k][1].record()
        self.barrier()
        w1 = time.time()
        return sum(a.elapsed_time(b) for a, b in ev), (w0, w1)


def _roofline(agent_steps_per_launch, launch_ms, bytes_per_agent_step, kernel, traffic=None, traffic_source=None):
    peak, which = _peaks()
    achieved = agent_steps_per_launch * bytes_per_agent_step / (launch_ms * 1e-3) / 1e9
    out = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
           "peak_source": which, "kernel": kernel, "bytes_per_agent_step": bytes_per_agent_step,
           "agent_steps_per_launch": agent_steps_per_launch, "launch_ms": launch_ms}
    if traffic is not None:
        out["traffic_source"] = traffic_source
        out["dram_gbs"] = traffic / (launch_ms * 1e-3) / 1e9  # real DRAM traffic per second of the launch
        out["dram_frac"] = out["dram_gbs"] / peak
    return out


def measure_traffic(args):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE engine-kernel launch (a 91-tick rollout), measured now by a
    child run under ncu (two metrics, one pass); falls back to the committed capture of this round."""
    fallback = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "engine_kernel_traffic.json")) as f:
            d = json.load(f)
            fallback = d.get("dram_bytes_per_launch"), "profiles/engine_kernel_traffic.json (" + d.get("how", "") + ")"
    except Exception:
        pass
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if args.no_ncu or not os.path.exists(ncu):
        return fallback
    try:
        with tempfile.TemporaryDirectory() as tmp:
            log = os.path.join(tmp, "t.csv")
            cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k",
                   "regex:lcs_engine_kernel", "--launch-skip", "3", "--launch-count", "2", "--csv", "--log-file", log,
                   sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "replay", "--kernel-only", "--steps", "1",
                   "--warmup", "1", "--rollouts-per-step", "1", "--scenarios", str(args.scenarios), "--no-ncu"]
            env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK=os.environ.get("LOCAL_RANK", "0"))
            subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=240, env=env, check=True)
            # launches alternate reset (1 tick) / rollout (91 ticks): two consecutive ones are captured, the rollout is
            # the one with the larger traffic
            per_launch = {}
            for line in open(log):
                parts = [x.strip('"') for x in line.strip().split('","')]
                if len(parts) > 3 and parts[-3].startswith("dram__bytes"):
                    v, unit = float(parts[-1].replace(",", "")), parts[-2]
                    per_launch[parts[0]] = per_launch.get(parts[0], 0.0) + v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            total = max(per_launch.values()) if per_launch else 0.0
            if total > 0:
                return total, "ncu child run of this bench (dram__bytes_read.sum + dram__bytes_write.sum of one 91-tick rollout launch)"
    except Exception:
        pass
    return fallback


_BATCH_CACHE: dict = {}


def bench_batch(ctx: Ctx, kind: str, steps: int, warmup: int, traffic=(None, None)):
    """replay / reactive / replan / rl_env on a BatchEngine; returns the workload's record"""
    import numpy as np

    from lcsim_b200 import native, shard, workload
    from lcsim_b200.batch import BatchEngine

    args, torch, dev = ctx.args, ctx.torch, ctx.dev
    S = args.scenarios if kind in ("replay", "reactive") else args.env_scenarios
    R = max(1, args.rollouts_per_step // (3 if kind != "replay" else 1))  # rollouts per step
    reactive = kind == "reactive"
    t_init = time.time()
    # (log-replay and reactive run the same batch: generated once per process, `init_s` of the second is its upload)
    key = (ctx.rank * S + (0 if kind in ("replay", "reactive") else 100000), S, args.agents, kind == "rl_env")
    if key not in _BATCH_CACHE:
        _BATCH_CACHE.clear()
        _BATCH_CACHE[key] = workload.synthetic_batch(key[0], S, args.agents, workers=max(1, min(32, (os.cpu_count() or 1) // ctx.world)),
                                                     with_centres=kind == "rl_env")
    sb = _BATCH_CACHE[key]
    if kind == "rl_env":
        sb, centres = sb
    be = BatchEngine(sb, reactive=reactive, with_metrics=not args.no_metrics, device=ctx.local, grid_cell=args.grid_cell)
    Ap = be.A_stride
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()  # noqa: E731
    stats_acc = torch.zeros(native.N_STATS, dtype=torch.int64, device=dev)
    kev = []  # CUDA events around the stepping launches alone (the roofline's launch duration)
    obs_bytes = 0.0
    extra = {}

    if kind == "replan":
        # Engine._plan_trajectory (engine.py:286-297): at the prepare of step 10, 90 (plan_init_step 10, plan_interval 80)
        # every dynamic ACTIVE agent receives a float32 (80, .) plan; random tensors stand in for the denoiser
        dyn = torch.from_numpy(np.ascontiguousarray(sb.dynamic[:, : be.A]).astype(np.bool_)).to(dev)
        g = torch.Generator(device=dev)
        g.manual_seed(7 + ctx.rank)
        plan_ticks = [k for k in range(N_TICKS) if k >= PLAN_INIT and (k - PLAN_INIT) % PLAN_INTERVAL == 0]

        def make_plans():
            sel = (dyn & (be.status == native.ACTIVE)).nonzero()  # one host sync, like reading the denoiser's batch
            n = int(sel.shape[0])
            pose = be.pose64[sel[:, 0], sel[:, 1]]
            h = (pose[:, 2:3] + 0.01 * torch.randn(n, PLAN_T, generator=g, device=dev, dtype=torch.float64).cumsum(1))
            speed = 12.0 * torch.rand(n, 1, generator=g, device=dev, dtype=torch.float64)
            d = torch.stack([h.cos(), h.sin()], dim=-1) * speed.unsqueeze(-1)
            xy = (pose[:, None, :2] + (d * 0.1).cumsum(1)).float()
            return sel[:, 0].int().contiguous(), sel[:, 1].int().contiguous(), xy.contiguous(), d.float().contiguous(), h.float().contiguous()

        def one_rollout(timed=False):
            be.reset()
            prev = 0
            for k in plan_ticks + [N_TICKS]:
                if k > prev:
                    if timed:
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                    be.rollout(k - prev)
                    if timed:
                        b.record()
                        kev.append((a, b))
                prev = k
                if k < N_TICKS:
                    be.set_ref_trajectory(*make_plans())
        extra["plans_per_rollout"] = len(plan_ticks)
    elif kind == "rl_env":
        # configs[2]: every scenario is one BicycleSingleEnv (gym_warpper.py:78-87, waymo_ppo.yml: reactive False,
        # total 90): per tick a policy action per env goes in, the tick runs, the ego-centric observation tensors are
        # built on the device (A10) and reward / terminated / truncated / route come back
        be.upload_polylines(centres)
        rng = np.random.default_rng(1 + ctx.rank)
        acts = torch.from_numpy(np.stack([rng.uniform(-0.3, 0.3, (N_TICKS, S)), rng.uniform(-0.2, 0.2, (N_TICKS, S))],
                                         axis=-1)).contiguous().to(dev)
        K_A, K_P = 32, 64
        P = int(be.P)
        obs_bytes_env_tick = 12.0 * P + 16.0 * (K_A + K_P)  # polyline-centre read + edge/feature writes
        extra.update(k_agents=K_A, k_polys=K_P, polyline_centres=P)

        # device-resident arm: lcsim_b200_env_step_device (actions and result records stay in HBM) with the observation
        # registered once (lcsim_b200_set_env_obs): per tick ONE call = tick items | metric items + reward kernel beside
        # the observation build
        be.set_env_obs(k_agents=K_A, k_polys=K_P, sort_by_distance=True)
        acts_k = [acts[k].contiguous() for k in range(90)]
        env_buf = [None]

        def one_rollout(timed=False):
            be.reset()
            if timed:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            for k in range(90):
                env_buf[0] = be.env_step_device(acts_k[k], env_buf[0], coef=WAYMO_PPO_REWARD)
            if timed:
                b.record()
                kev.append((a, b))
    else:
        def one_rollout(timed=False):
            be.reset()
            if timed:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            be.rollout(N_TICKS)
            if timed:
                b.record()
                kev.append((a, b))
    init_s = time.time() - t_init

    # ------------------------------------------------------------- device-resident arm ("value")
    def step_dev():
        for _ in range(R):
            one_rollout(timed=True)
            stats_acc.add_(be.episode_stats())

    one_rollout()
    kev.clear()
    stats_acc.zero_()
    launches0 = be.launches
    for _ in range(warmup):
        step_dev()
    kev.clear()
    stats_acc.zero_()
    launches0 = be.launches
    ms, win = ctx.timed(step_dev, steps, 0)
    launches = be.launches - launches0
    total = shard.reduce_episode_stats(stats_acc)  # the one collective of the path: NCCL all-reduce over NVLink
    ms_max, _ = ctx.max_sum(ms, 0.0)
    agent_steps_all, agent_steps_rank = int(total[0].item()), int(stats_acc[0].item())
    value = agent_steps_all / (ms_max * 1e-3)
    n_step_launches = len(kev)
    launch_ms = sum(a.elapsed_time(b) for a, b in kev) / max(n_step_launches, 1)
    if kind == "rl_env":
        obs_bytes = obs_bytes_env_tick * S * 90 * R * steps / max(agent_steps_rank, 1)
    bpa = b_alg(fp64_master=kind in ("reactive", "rl_env", "replan"), obs_bytes_per_agent_step=obs_bytes)
    rec = {
        "value": value, "unit": "agent-steps/s", "steps": steps, "warmup": warmup, "ms_per_step": ms_max / steps,
        "rollouts_per_step": R, "scenarios_per_gpu": S, "agent_steps_per_rollout": agent_steps_all // max(steps * R * ctx.world, 1),
        "gpu_launches": int(launches), "init_s": round(init_s, 2), "clocks": ctx.sampler.window(*win),
        "roofline": _roofline(agent_steps_rank / max(n_step_launches, 1), launch_ms, bpa,
                              "lcs_engine_kernel (one launch = %s)" % ("91 ticks" if kind in ("replay", "reactive") else
                                                                       "the ticks between two re-plans" if kind == "replan" else
                                                                       "90 x (tick items, metric items, obs and rewards kernels)"),
                              *(traffic if kind == "replay" else (None, None))),
        "episode": shard.rates(total), **extra,
    }

    # ------------------------------------------------------------- end-to-end arm ("e2e"): host buffers
    if not args.kernel_only:
        if kind in ("replay", "reactive"):
            # lcsim_b200_rollout_host: `engine.reset(); for _ in range(91): engine.step()` for every scenario and the
            # read-back of what a caller of the reference reads, ONE call with pinned host buffers
            r_out = dict(tick_log=pin((N_TICKS, S, 4), torch.int32), pose32=pin((S, Ap, 4), torch.float32),
                         status=pin((S, Ap), torch.int32), coll_count=pin((S, Ap), torch.int32),
                         nearest_edge=pin((S, Ap), torch.int32), offroad=pin((S, Ap), torch.uint8), stats=pin((8,), torch.int64))
            r_mask = torch.ones(S, dtype=torch.uint8).pin_memory()
            h2d, d2h = r_mask.numel(), sum(v.numel() * v.element_size() for v in r_out.values())
            cnt = [0]

            def step_e2e():
                for _ in range(R):
                    be.rollout_host(N_TICKS, r_out, reset_mask=r_mask)  # returns after the stream sync
                    cnt[0] += int(r_out["stats"][0])
            what = ("lcsim_b200_rollout_host, one call per rollout with pinned host buffers: H2D reset mask, reset + 91 ticks, "
                    "D2H of the per-tick log [91][S][4], pose32, status, collision counts, nearest edge, off-road flags, "
                    "statistics; stream sync before the host reads them")
        elif kind == "rl_env":
            # lcsim_b200_env_step == BicycleSingleEnv.step for all S envs: every tick the normalised ADS actions go host ->
            # device from pinned memory, tick + observation build + reward kernel run, reward / terms / route / flags come
            # back device -> host and the host reads them (stream sync) before the next tick, like an RL loop does
            h_acts = acts.cpu().pin_memory()
            h_out = torch.empty(S * native.ENV_OUT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
            out_np = h_out.numpy().view(native.ENV_OUT_DTYPE)
            h_stats = torch.empty(native.N_STATS, dtype=torch.int64).pin_memory()
            h2d, d2h = 90 * S * 16, 90 * h_out.numel() + 64
            cnt, ret = [0], [0.0]

            be.set_env_obs(k_agents=K_A, k_polys=K_P, sort_by_distance=True)  # env_step builds it behind every tick
            h_act_ptrs = [h_acts[k] for k in range(90)]

            def step_e2e():
                for _ in range(R):
                    be.reset()
                    for k in range(90):
                        be.env_step(h_act_ptrs[k], h_out, coef=WAYMO_PPO_REWARD, sync=True)  # returns after the stream sync
                        ret[0] += float(out_np["reward"][0])  # the host consumes the result
                    h_stats.copy_(be.episode_stats(), non_blocking=True)
                    torch.cuda.current_stream(dev).synchronize()
                    cnt[0] += int(h_stats[0])
            what = ("per tick ONE call lcsim_b200_env_step with the observation registered by lcsim_b200_set_env_obs: pinned "
                    "H2D of every env's ADS action, tick, reward kernel next to the observation build (device tensors for "
                    "the encoder), reward/terms/route/flags written by the reward kernel straight into the pinned host "
                    "buffer (224 B per env across PCIe), stream sync, host read; per rollout D2H of the statistics")
        else:
            # plans arrive in pinned HOST memory (the reference hands numpy arrays to set_ref_trajectory), go H2D, and the
            # rollout ends with the read-back of rollout_host
            r_out = dict(pose32=pin((S, Ap, 4), torch.float32), status=pin((S, Ap), torch.int32), coll_count=pin((S, Ap), torch.int32),
                         nearest_edge=pin((S, Ap), torch.int32), offroad=pin((S, Ap), torch.uint8), stats=pin((8,), torch.int64))
            d2h = sum(v.numel() * v.element_size() for v in r_out.values())
            h2d, cnt = 0, [0]
            n_max = S * be.A
            plan_pins = [pin((n_max,), torch.int32), pin((n_max,), torch.int32), pin((n_max, PLAN_T, 2), torch.float32),
                         pin((n_max, PLAN_T, 2), torch.float32), pin((n_max, PLAN_T), torch.float32)]

            def step_e2e():
                nonlocal h2d
                for _ in range(R):
                    be.reset()
                    prev, h2d = 0, 0
                    for k in plan_ticks + [N_TICKS]:
                        if k > prev:
                            be.rollout(k - prev)
                        prev = k
                        if k < N_TICKS:
                            # the denoiser's output arrives as HOST arrays (reused pinned buffers), then goes H2D
                            plans = make_plans()
                            n = int(plans[0].shape[0])
                            host = [buf[:n] for buf in plan_pins]
                            for hb, t in zip(host, plans):
                                hb.copy_(t, non_blocking=True)
                            torch.cuda.current_stream(dev).synchronize()  # the host holds the plans now
                            h2d += sum(t.numel() * t.element_size() for t in host)
                            be.set_ref_trajectory(*[t.to(dev, non_blocking=True) for t in host])
                    for name, buf in r_out.items():
                        if name != "stats":
                            buf.copy_(getattr(be, name) if name != "pose32" else be.pose32, non_blocking=True) if buf.shape[1] == be.A else None
                    r_out["stats"].copy_(be.episode_stats(), non_blocking=True)
                    torch.cuda.current_stream(dev).synchronize()
                    cnt[0] += int(r_out["stats"][0])
            what = ("per rollout: reset, rollout to each plan tick, plans (scenario, agent, xy, vel, heading float32) from "
                    "pinned host memory H2D + lcsim_b200_set_ref_trajectory, rollout to the end, D2H of the statistics; sync")
        step_e2e()
        cnt[0] = 0
        ms_e, _ = ctx.timed(step_e2e, steps, min(warmup, 2), flush=False)
        ms_e_max, n_all = ctx.max_sum(ms_e, float(cnt[0]) * steps / (steps + min(warmup, 2)))
        rec["e2e"] = {"value": n_all / (ms_e_max * 1e-3), "unit": "agent-steps/s", "h2d_bytes_per_step": int(h2d * R),
                      "d2h_bytes_per_step": int(d2h * R), "what": what}
    be.close()
    return rec, sb


def bench_city(ctx: Ctx, steps: int, warmup: int):
    """configs[3]: one MOSS-shaped city, 200k vehicles on one road network, LaneIDM + spatial-hash collision +
    off-road every tick.  One scenario does not shard: every rank runs a replica."""
    from lcsim_b200 import city
    from lcsim_b200.city_engine import CityEngine

    args, torch, dev = ctx.args, ctx.torch, ctx.dev
    t0 = time.time()
    st = city.pack_city(city.make_city(seed=42 + ctx.rank, grid=args.city_grid, n_agents=args.city_agents, depart_window=10.0, max_hops=6))
    ce = CityEngine(st, device=ctx.local)
    init_s = time.time() - t0
    TICKS = args.city_ticks
    stats_acc = torch.zeros(8, dtype=torch.int64, device=dev)

    def step_dev():
        ce.reset()
        ce.step(TICKS)
        stats_acc.add_(ce.stats)

    step_dev()
    stats_acc.zero_()
    l0 = ce.launches
    for _ in range(warmup):
        step_dev()
    stats_acc.zero_()
    l0 = ce.launches
    ms, win = ctx.timed(step_dev, steps, 0)
    ms_max, n_all = ctx.max_sum(ms, float(stats_acc[0].item()))
    launches = ce.launches - l0
    per_tick_ms = ms / (steps * TICKS)
    rec = {"value": n_all / (ms_max * 1e-3), "unit": "agent-steps/s", "steps": steps, "warmup": warmup, "ms_per_step": ms_max / steps,
           "ticks_per_step": TICKS, "us_per_tick": 1e3 * per_tick_ms, "vehicles": int(st.n_agents), "lanes": int(st.n_lanes),
           "gpu_launches": int(launches), "launches_per_tick": launches / (steps * TICKS), "init_s": round(init_s, 1),
           "scaling": "replicas only (one scenario per GPU)", "clocks": ctx.sampler.window(*win),
           "roofline": _roofline(float(stats_acc[0].item()) / (steps * TICKS), per_tick_ms, B_CITY,
                                 "city tick (CUDA graph of the lcc_* kernels); launch = one tick")}
    if not args.kernel_only:
        pin = lambda t: torch.empty(t.shape, dtype=t.dtype).pin_memory()  # noqa: E731
        views = [ce.xy, ce.heading, ce.v, ce.status, ce.coll_count, ce.offroad, ce.stats]
        host = [pin(v) for v in views]
        cnt = [0]

        def step_e2e():
            ce.reset()
            ce.step(TICKS)
            for h, v in zip(host, views):
                h.copy_(v, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            cnt[0] += int(host[-1][0])

        step_e2e()
        cnt[0] = 0
        ms_e, _ = ctx.timed(step_e2e, steps, 0, flush=False)
        ms_e_max, n_e = ctx.max_sum(ms_e, float(cnt[0]))
        rec["e2e"] = {"value": n_e / (ms_e_max * 1e-3), "unit": "agent-steps/s", "h2d_bytes_per_step": 0,
                      "d2h_bytes_per_step": int(sum(h.numel() * h.element_size() for h in host)),
                      "what": "reset + %d ticks (lcsim_b200_city_step) + D2H of xy, heading, v, status, collision counts, "
                              "off-road flags, statistics into pinned host memory; stream sync" % TICKS}
    ce.close()
    return rec


def run_b200(args):
    ctx = Ctx(args)
    kinds = WORKLOADS if args.workload == "all" else (args.workload,)
    head_kind = kinds[0]
    t_start = time.time()
    traffic = (None, None)
    if ctx.rank == 0 and head_kind == "replay" and not args.kernel_only:
        traffic = measure_traffic(args)
    records, sb_head = {}, None
    for kind in kinds:
        sub = kind != head_kind
        steps = args.steps if not sub else max(3, args.steps // 2)
        warmup = args.warmup if not sub else min(args.warmup, 3)
        if kind == "city":
            records[kind] = bench_city(ctx, steps, warmup)
        else:
            records[kind], sb = bench_batch(ctx, kind, steps, warmup, traffic)
            if not sub:
                sb_head = sb
    ctx.sampler.stop()
    if ctx.rank == 0:
        head = records[head_kind]
        cores = os.cpu_count() or 1
        out = {
            "metric": METRIC, "value": head["value"], "unit": "agent-steps/s", "n_gpus": ctx.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": _workload_text(args) if head_kind == "replay" else head_kind,
                       "scenarios_per_gpu": head.get("scenarios_per_gpu"), "agents": args.agents, "ticks": N_TICKS,
                       "mode": {"replay": "log-replay", "reactive": "reactive (TrajIDM + bicycle)"}.get(head_kind, head_kind),
                       "l2": "inputs 5.6 GB/GPU > 126 MB L2; L2 flushed between timed steps",
                       "exact_shortcuts": "replay records (cursor of a replayed waypoint from a canonical-index table instead of the "
                                          "trajectory scan; LCSIM_B200_RECORDS=0 disables), edge search skipped for agents whose position "
                                          "did not change, per-cell candidate lists for the nearest road-edge point "
                                          "(LCSIM_B200_CAND_CELL=0 disables), windowed cursor scan with safe radii; results "
                                          "bit-identical either way (tests/test_parity.py)"},
            "clocks": head["clocks"], "e2e": head.get("e2e", {"value": None, "unit": "agent-steps/s", "h2d_bytes_per_step": 0,
                                                                 "d2h_bytes_per_step": 0}),
            "gpu_launches": sum(r["gpu_launches"] for r in records.values()),
            "gpu_launches_headline": head["gpu_launches"], "roofline": head["roofline"],
            "agent_steps_per_rollout": head.get("agent_steps_per_rollout"), "rollouts_per_step": head.get("rollouts_per_step"),
            "init_s": head.get("init_s"), "episode": head.get("episode"),
        }
        for kind in kinds[1:]:
            out[kind] = records[kind]
        if not args.kernel_only and head_kind in ("replay", "reactive") and sb_head is not None:
            n_cpu = args.cpu_scenarios
            cpu_steps, cpu_secs = cpu_rollouts(sb_head.scenario_slice(0, n_cpu), cores, reactive=head_kind == "reactive")
            out["cpu_baseline"] = {"value": cpu_steps / cpu_secs, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                                   "sample": f"first {n_cpu} scenarios of rank 0's batch, one 91-tick rollout with metrics, "
                                             f"oracle/lcsim_oracle.c on {cores} threads"}
            if "reactive" in records and head_kind == "replay":
                c2, s2 = cpu_rollouts(sb_head.scenario_slice(0, n_cpu), cores, reactive=True)
                out["reactive"]["cpu_baseline"] = {"value": c2 / s2, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                                                   "sample": f"same {n_cpu} scenarios, reactive"}
        out["wall_s"] = round(time.time() - t_start, 1)
        print_json(out)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


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
    ap.add_argument("--workload", default="all", choices=("all",) + WORKLOADS,
                    help="all: log-replay headline + a sub-record per other BASELINE config; or one workload as the headline")
    ap.add_argument("--rollouts-per-step", type=int, default=50, help="rollouts of the batch per timed step (headline; a third "
                    "for the other batch workloads)")
    ap.add_argument("--scenarios", type=int, default=1024, help="scenarios per GPU (configs[1])")
    ap.add_argument("--env-scenarios", type=int, default=512, help="envs / scenarios per GPU of rl_env (configs[2]: 4096 on 8 "
                    "GPUs) and replan (configs[4]: 512)")
    ap.add_argument("--agents", type=int, default=128)
    ap.add_argument("--city-agents", type=int, default=200000)
    ap.add_argument("--city-grid", type=int, default=70)
    ap.add_argument("--city-ticks", type=int, default=300, help="ticks per timed step of the city workload")
    ap.add_argument("--grid-cell", type=float, default=0.0, help="experiment: road-edge grid cell (m); 0 = library default")
    ap.add_argument("--no-metrics", action="store_true", help="experiment: skip the collision/off-road metric items")
    ap.add_argument("--kernel-only", action="store_true", help="skip the e2e and CPU legs (profiling runs)")
    ap.add_argument("--no-ncu", action="store_true", help="do not measure DRAM traffic with an ncu child run")
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
