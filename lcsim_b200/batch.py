"""Batched engine: S independent scenarios stepped by one CUDA launch per tick.

This is the host-side owner of one ``lcsim_b200_engine`` handle (include/lcsim_b200.h).  It mirrors,
for a whole batch at once, the reference's ``Engine.reset / step`` (lcsim/engine/engine.py:132-158),
``AgentManager.check_collision_all / get_collision_rate / get_offroad_rate``
(lcsim/entity/agent/manager.py:313-345), ``Engine._compute_rewards`` (engine.py:301-357),
``Agent.get_route / set_ref_trajectory`` (lcsim/entity/agent/agent.py:462-480).  The per-scenario
facade with the reference's class names lives in ``lcsim_b200/engine.py``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import native
from .packer import StaticBatch


class BatchEngine:
    def __init__(self, static: StaticBatch, reactive: bool = False, no_static: bool = False,
                 allow_new_obj: bool = True, with_metrics: bool = True, device: int = 0, grid_cell: float = 0.0,
                 backend=None):
        self.backend = backend if backend is not None else native.TorchCudaBackend(device)
        self.lib = self.backend.lib
        self.static = static
        self.S, self.A, self.T, self.E = static.S, static.A, static.T, static.E
        self.reactive, self.no_static, self.allow_new_obj = bool(reactive), bool(no_static), bool(allow_new_obj)
        self.with_metrics = bool(with_metrics)
        d = native.BatchDesc()
        d.abi_version = native.ABI_VERSION
        d.S, d.A, d.T, d.E = static.S, static.A, static.T, static.E
        d.start, d.interval, d.total_steps = static.start, static.interval, static.total_steps
        d.reactive, d.no_static, d.allow_new_obj = int(reactive), int(no_static), int(allow_new_obj)
        d.device = getattr(self.backend, "device_index", 0)
        d.with_metrics = int(with_metrics)
        d.grid_cell = float(grid_cell)
        self._h = C.c_void_p()
        rc = self.lib.lcsim_b200_create(C.byref(d), C.byref(self._h))
        if rc != 0:
            msg = self.lib.lcsim_b200_last_error(self._h).decode() if self._h else "create failed"
            self._destroy()
            raise native.NativeLibraryError(f"lcsim_b200_create: {rc}: {msg}")
        sh = native.StaticHost()
        keep = []
        for name, dt in native.STATIC_FIELDS:
            val = getattr(static, name, None)
            if val is None:  # optional arrays (edge_poly_id of batches packed before ABI 2)
                setattr(sh, name, None)
                continue
            arr = np.ascontiguousarray(val, dtype=dt)
            keep.append(arr)
            setattr(sh, name, arr.ctypes.data)
        self._check(self.lib.lcsim_b200_upload_static(self._h, C.byref(sh)), "upload_static")
        del keep
        v = native.StateViews()
        self._check(self.lib.lcsim_b200_views(self._h, C.byref(v)), "views")
        self.A_stride = int(v.A_stride)
        S, Ap, T, H = self.S, self.A_stride, self.T, native.H_STEPS
        shapes = {"tick": (S,), "n_active": (S,), "active": (S, Ap), "status": (S, Ap), "cursor": (S, Ap),
                  "ring_count": (S, Ap), "ring_head": (S, Ap), "pose64": (S, Ap, 4), "pose32": (S, Ap, 4),
                  "ring": (S, Ap, H, 5), "lead_valid": (S, Ap), "lead_dis": (S, Ap), "lead_v": (S, Ap),
                  "coll_count": (S, Ap), "nearest_edge": (S, Ap), "offroad": (S, Ap), "traj_len": (S, Ap),
                  "traj_xy": (S, Ap, T, 2), "traj_vel": (S, Ap, T, 2), "traj_h": (S, Ap, T),
                  "stats": (S, native.N_STATS), "origin": (S, 2)}
        # zero-copy views of the engine's buffers; per-agent arrays are sliced to the logical A
        for name, typestr in native.VIEW_FIELDS:
            t = self.backend.view(getattr(v, name), shapes[name], typestr)
            if len(shapes[name]) >= 2 and shapes[name][1] == Ap and name not in ("stats", "origin"):
                t = t[:, : self.A]
            setattr(self, name, t)
        self._stats_out = self.backend.empty((native.N_STATS,), np.int64)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise native.NativeLibraryError(f"lcsim_b200_{what}: {rc}: {self.lib.lcsim_b200_last_error(self._h).decode()}")

    def _destroy(self):
        if getattr(self, "_h", None):
            self.lib.lcsim_b200_destroy(self._h)
            self._h = None

    def close(self):
        self._destroy()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.lcsim_b200_launch_count(self._h))

    def sync(self):
        self.backend.sync()

    # ------------------------------------------------------------------ Engine.reset / step
    def reset(self, mask=None):
        """Engine.reset for the scenarios selected by ``mask`` ((S,) bool/uint8; None = all)."""
        if mask is None:
            self._check(self.lib.lcsim_b200_reset(self._h, None, self.backend.stream()), "reset")
            return
        m = self.backend.to_device(mask, np.uint8)
        self._check(self.lib.lcsim_b200_reset(self._h, self.backend.ptr(m), self.backend.stream()), "reset")
        self._keep = m

    def step(self, act_agent=None, act_type=None, act_data=None, mask=None):
        """Engine.step(actions).  act_agent/act_type: (S,K) int32, act_data: (S,K,5) float64 (device tensors or
        arrays); None lets every agent use its own policy.  mask: (S,) bool/uint8, only those scenarios advance
        (None = all)."""
        m = None if mask is None else self.backend.to_device(mask, np.uint8)
        mp = None if m is None else self.backend.ptr(m)
        if act_agent is None:
            self._check(self.lib.lcsim_b200_step_masked(self._h, mp, None, None, None, 0, self.backend.stream()), "step")
            self._keep = m
            return
        aa = self.backend.to_device(act_agent, np.int32)
        at = self.backend.to_device(act_type, np.int32)
        ad = self.backend.to_device(act_data, np.float64)
        K = int(aa.shape[1])
        assert tuple(aa.shape) == tuple(at.shape) == (self.S, K) and tuple(ad.shape) == (self.S, K, 5)
        self._check(self.lib.lcsim_b200_step_masked(self._h, mp, self.backend.ptr(aa), self.backend.ptr(at),
                                                    self.backend.ptr(ad), K, self.backend.stream()), "step")
        self._keep = (aa, at, ad, m)  # the launch is asynchronous: keep the buffers alive until the next call

    def rollout(self, n_ticks: int, log: bool = False):
        """``for _ in range(n_ticks): engine.step()`` for the whole batch in one launch.  With ``log`` returns the
        (n_ticks, S, 4) int32 device tensor of per-tick [agents updated, active after, collision sum, off-road count]."""
        if not log:
            self._check(self.lib.lcsim_b200_rollout(self._h, int(n_ticks), self.backend.stream()), "rollout")
            return None
        out = self.backend.empty((int(n_ticks), self.S, 4), np.int32)
        self._check(self.lib.lcsim_b200_rollout_log(self._h, int(n_ticks), self.backend.ptr(out), self.backend.stream()),
                    "rollout_log")
        return out

    def rollout_host(self, n_ticks: int, out: dict, reset_mask=None, sync: bool = True):
        """reset + n_ticks steps + read-back into HOST buffers in one call (lcsim_b200_rollout_host).  ``out`` maps the
        field names of lcsim_b200_rollout_out to host arrays / pinned tensors: tick_log (n_ticks,S,4) i32, pose32
        (S,A_stride,4) f32, status / coll_count / nearest_edge (S,A_stride) i32, offroad (S,A_stride) u8, stats (8,) i64;
        ``reset_mask`` a host (S,) uint8 array or None."""
        def host_ptr(x):
            return x.ctypes.data if isinstance(x, np.ndarray) else int(x.data_ptr())

        ro = native.RolloutOut()
        for k, v in out.items():
            setattr(ro, k, host_ptr(v))
        self._check(self.lib.lcsim_b200_rollout_host(self._h, None if reset_mask is None else host_ptr(reset_mask), int(n_ticks),
                                                     C.byref(ro), int(sync), self.backend.stream()), "rollout_host")
        self._keep = (out, reset_mask)

    # ------------------------------------------------------------------ diffusion-action ingest
    def set_ref_trajectory(self, scenario, agent, xy, vel, heading):
        """Agent.set_ref_trajectory for n (scenario, agent) pairs: xy/vel (n,T,2), heading (n,T) float32."""
        sc = self.backend.to_device(scenario, np.int32)
        ag = self.backend.to_device(agent, np.int32)
        xy = self.backend.to_device(xy, np.float32)
        vel = self.backend.to_device(vel, np.float32)
        hd = self.backend.to_device(heading, np.float32)
        n, T = int(hd.shape[0]), int(hd.shape[1])
        assert tuple(xy.shape) == tuple(vel.shape) == (n, T, 2) and tuple(sc.shape) == tuple(ag.shape) == (n,)
        self._check(self.lib.lcsim_b200_set_ref_trajectory(
            self._h, self.backend.ptr(sc), self.backend.ptr(ag), self.backend.ptr(xy), self.backend.ptr(vel),
            self.backend.ptr(hd), n, T, self.backend.stream()), "set_ref_trajectory")
        self._keep = (sc, ag, xy, vel, hd)

    # ------------------------------------------------------------------ rewards / route
    def _coef6(self, coef: Optional[dict]):
        """host (6,) float64 array of the reward coefficients (cached per dict: the per-tick calls pass the same one)"""
        if coef is None:
            return None
        key = tuple(float(coef.get(k, 0.0)) for k in native.REWARD_KEYS)
        hit = getattr(self, "_coef_cache", None)
        if hit is None or hit[0] != key:
            hit = self._coef_cache = (key, np.asarray(key, np.float64))
        return hit[1]

    def rewards(self, agent=None, coef: Optional[dict] = None, out: Optional[dict] = None):
        """Engine._compute_rewards + get_route for one agent per scenario (None = the ADS).
        Returns dict(terms (S,6), terminated (S,), truncated (S,), route (S,10,2), reward (S,)); ``out`` = a dict returned
        by an earlier call is written in place (a per-tick loop then allocates nothing)."""
        b = self.backend
        if out is None:
            out = dict(terms=b.empty((self.S, native.N_REWARD_TERMS), np.float64), terminated=b.empty((self.S,), np.uint8),
                       truncated=b.empty((self.S,), np.uint8), route=b.empty((self.S, native.ROUTE_LEN, 2), np.float64),
                       reward=b.empty((self.S,), np.float64))
        ag = None if agent is None else b.to_device(agent, np.int32)
        cf = self._coef6(coef)
        self._check(self.lib.lcsim_b200_rewards(
            self._h, None if ag is None else b.ptr(ag), None if cf is None else cf.ctypes.data,
            b.ptr(out["reward"]), b.ptr(out["terms"]), b.ptr(out["terminated"]), b.ptr(out["truncated"]),
            b.ptr(out["route"]), b.stream()), "rewards")
        self._keep = ag
        return out

    # ------------------------------------------------------------------ batched env step with host buffers
    def env_step(self, actions, out, coef: Optional[dict] = None, sync: bool = True):
        """BicycleSingleEnv.step for every scenario at once (lcsim/envs/gym_warpper.py:78-87).

        actions: (S,2) float64 HOST array/tensor of normalised (acc, steer) for each scenario's ADS (None: own policy);
        out: HOST buffer of S records of native.ENV_OUT_DTYPE (numpy structured array, or a pinned uint8 tensor of
        S*224 bytes).  H2D copy, tick, reward kernel and D2H copy are enqueued on the current stream."""
        def host_ptr(x):
            return x.ctypes.data if isinstance(x, np.ndarray) else int(x.data_ptr())

        cf = self._coef6(coef)
        self._check(self.lib.lcsim_b200_env_step(self._h, None if actions is None else host_ptr(actions),
                                                 None if cf is None else cf.ctypes.data, host_ptr(out), int(sync),
                                                 self.backend.stream()), "env_step")
        self._keep = (actions, out)

    def env_step_device(self, actions, out=None, coef: Optional[dict] = None):
        """:meth:`env_step` with DEVICE buffers (lcsim_b200_env_step_device): actions (S,2) float64 device tensor or None,
        out a device uint8 tensor of S * ENV_OUT_DTYPE.itemsize bytes (allocated when None; reuse it).  Nothing is copied
        or waited for; a registered observation (:meth:`set_env_obs`) is built beside the metric items."""
        b = self.backend
        if out is None:
            out = b.empty((self.S * native.ENV_OUT_DTYPE.itemsize,), np.uint8)
        a = None if actions is None else b.to_device(actions, np.float64)
        self._check(self.lib.lcsim_b200_env_step_device(self._h, None if a is None else b.ptr(a), self._coef6_ptr(coef),
                                                        b.ptr(out), b.stream()), "env_step_device")
        self._keep = (a, out)
        return out

    def _coef6_ptr(self, coef):
        cf = self._coef6(coef)
        return None if cf is None else cf.ctypes.data

    # ------------------------------------------------------------------ observation builder
    def upload_polylines(self, centres):
        """centres: list of (P_s,3) arrays (x, y, heading) per scenario — packer.polyline_centres()."""
        assert len(centres) == self.S
        P = max(1, max(len(c) for c in centres))
        arr = np.zeros((self.S, P, 3))
        n = np.zeros(self.S, np.int32)
        for i, c in enumerate(centres):
            n[i] = len(c)
            arr[i, : len(c)] = c
        self._check(self.lib.lcsim_b200_upload_polylines(self._h, n.ctypes.data, arr.ctypes.data, P), "upload_polylines")
        self.P = P

    def _obs_desc(self, k_agents, k_polys, radius, sort_by_distance, ego, with_polys, with_history, out):
        b = self.backend
        if out is None:
            out = dict(nbr_agent=b.empty((self.S, k_agents), np.int32), nbr_agent_feat=b.empty((self.S, k_agents, 3), np.float32),
                       n_nbr_agent=b.empty((self.S,), np.int32))
            if with_polys:
                out.update(nbr_poly=b.empty((self.S, k_polys), np.int32), nbr_poly_feat=b.empty((self.S, k_polys, 3), np.float32),
                           n_nbr_poly=b.empty((self.S,), np.int32))
            if with_history:
                out.update(history=b.empty((self.S, self.A_stride, native.H_STEPS, 6), np.float32),
                           history_agent=b.empty((self.S, self.A_stride), np.int32))
        d = native.ObsDesc()
        eg = None if ego is None else b.to_device(ego, np.int32)
        d.ego = None if eg is None else b.ptr(eg)
        d.radius, d.k_agents, d.k_polys, d.sort_by_distance = float(radius), int(k_agents), int(k_polys), int(sort_by_distance)
        for k, v in out.items():
            setattr(d, k, b.ptr(v))
        return d, out, eg

    def build_obs(self, k_agents: int = 32, k_polys: int = 64, radius: float = 50.0, sort_by_distance: bool = False,
                  ego=None, with_polys: bool = True, with_history: bool = True, out: Optional[dict] = None):
        """Ego-centric neighbour gather + history stack (include/lcsim_b200.h lcsim_b200_obs_desc).  Returns a dict of
        device tensors written in place by one launch; ``out`` = the dict of an earlier call with the same arguments is
        reused (nothing is allocated, the descriptor is kept)."""
        b = self.backend
        key = (k_agents, k_polys, radius, sort_by_distance, with_polys, with_history)
        hit = getattr(self, "_obs_cache", None)
        if out is not None and ego is None and hit is not None and hit[0] == key and hit[2] is out:
            d, eg = hit[1], None
        else:
            d, out, eg = self._obs_desc(k_agents, k_polys, radius, sort_by_distance, ego, with_polys, with_history, out)
            if ego is None:
                self._obs_cache = (key, d, out)
        self._check(self.lib.lcsim_b200_build_obs(self._h, C.byref(d), b.stream()), "build_obs")
        self._keep = eg
        return out

    def set_env_obs(self, k_agents: int = 32, k_polys: int = 64, radius: float = 50.0, sort_by_distance: bool = False,
                    ego=None, with_polys: bool = True, with_history: bool = True, enable: bool = True):
        """Registers the observation that every later :meth:`env_step` builds behind its tick (BicycleSingleEnv.step
        returns the observation together with the reward, gym_warpper.py:78-92).  Returns the dict of device tensors the
        engine keeps writing; ``enable=False`` removes the registration."""
        if not enable:
            self._check(self.lib.lcsim_b200_set_env_obs(self._h, None), "set_env_obs")
            self._env_obs = None
            return None
        d, out, eg = self._obs_desc(k_agents, k_polys, radius, sort_by_distance, ego, with_polys, with_history, None)
        self._check(self.lib.lcsim_b200_set_env_obs(self._h, C.byref(d)), "set_env_obs")
        self._env_obs = (out, eg, d)  # the engine holds raw pointers into these
        return out

    def build_env_obs(self):
        """Builds the registered observation now (after a reset, where no env_step ran); returns its tensors."""
        reg = getattr(self, "_env_obs", None)
        if reg is None:
            raise native.NativeLibraryError("build_env_obs: no observation registered (set_env_obs)")
        self._check(self.lib.lcsim_b200_build_obs(self._h, C.byref(reg[2]), self.backend.stream()), "build_obs")
        return reg[0]

    def build_encoder_inputs(self, a2a_radius: float = 50.0, pl2a_radius: float = 50.0, time_span: int = 10,
                             cap_a2a: Optional[int] = None, cap_pl2a: Optional[int] = None, with_polys: bool = True):
        """Encoder inputs of ALL listed agents (include/lcsim_b200.h lcsim_b200_encoder_desc): x_a, temporal edges and
        the a2a / pl2a edge lists in CSR form with the reference's relation features.  Returns a dict of device tensors
        written by one launch."""
        b = self.backend
        S, Ap, H = self.S, self.A_stride, native.H_STEPS
        cap_a2a = Ap * 32 if cap_a2a is None else int(cap_a2a)
        P = int(getattr(self, "P", 0)) if with_polys else 0
        cap_pl2a = max(1, P) * 16 if cap_pl2a is None else int(cap_pl2a)
        out = dict(agent=b.empty((S, Ap), np.int32), x_a=b.empty((S, Ap, H, 6), np.float32),
                   r_t=b.empty((S, Ap, H - 1, 4), np.float32), r_t_valid=b.empty((S, Ap, H - 1), np.uint8),
                   a2a_offset=b.empty((S, Ap + 1), np.int32), a2a_src=b.empty((S, cap_a2a), np.int32),
                   a2a_feat=b.empty((S, cap_a2a, 3), np.float32), n_a2a=b.empty((S,), np.int32))
        if with_polys and P:
            out.update(pl2a_offset=b.empty((S, P + 1), np.int32), pl2a_agent=b.empty((S, cap_pl2a), np.int32),
                       pl2a_feat=b.empty((S, cap_pl2a, 3), np.float32), n_pl2a=b.empty((S,), np.int32))
        d = native.EncoderDesc()
        d.a2a_radius, d.pl2a_radius, d.time_span = float(a2a_radius), float(pl2a_radius), int(time_span)
        d.cap_a2a, d.cap_pl2a = cap_a2a, cap_pl2a
        for k, v in out.items():
            setattr(d, k, b.ptr(v))
        self._check(self.lib.lcsim_b200_build_encoder_inputs(self._h, C.byref(d), b.stream()), "build_encoder_inputs")
        return out

    # ------------------------------------------------------------------ lane-centreline projection
    def upload_centrelines(self, points):
        """points: per scenario a tuple (xy_theta (N_s,3) float32, ids (N_s,) int32) — packer.centreline_points()."""
        assert len(points) == self.S
        N = max(1, max(len(p[0]) for p in points))
        arr = np.zeros((self.S, N, 3), np.float32)
        ids = np.zeros((self.S, N), np.int32)
        n = np.zeros(self.S, np.int32)
        for i, (xyt, pid) in enumerate(points):
            n[i] = len(xyt)
            arr[i, : len(xyt)] = xyt
            ids[i, : len(xyt)] = pid
        self._check(self.lib.lcsim_b200_upload_centrelines(self._h, n.ctypes.data, arr.ctypes.data, ids.ctypes.data, N),
                    "upload_centrelines")

    def project_lanes(self, top_k: int = 64):
        """Nearest lane-centreline point / lane id / OnLane distance / top-k heading difference of every active agent
        (include/lcsim_b200.h lcsim_b200_lane_desc).  Returns a dict of (S, A) device tensors."""
        b = self.backend
        out = dict(lane_pt=b.empty((self.S, self.A_stride), np.int32), lane_id=b.empty((self.S, self.A_stride), np.int32),
                   lane_dist=b.empty((self.S, self.A_stride), np.float32),
                   onlane_dist=b.empty((self.S, self.A_stride), np.float32),
                   heading_diff=b.empty((self.S, self.A_stride), np.float32))
        d = native.LaneDesc()
        d.top_k = int(top_k)
        for k, v in out.items():
            setattr(d, k, b.ptr(v))
        self._check(self.lib.lcsim_b200_project_lanes(self._h, C.byref(d), b.stream()), "project_lanes")
        return {k: v[:, : self.A] for k, v in out.items()}

    # ------------------------------------------------------------------ guidance costs / plan metrics
    def guidance_cost(self, which: str, xy, radius: Optional[float] = None, eps: float = 1e-6, with_sd: bool = False):
        """Repeller.loss / NoOffRoad.loss (motion_diff/modules/guidance.py:17-85) of the plans xy (S,A,T,2) float32 with
        the analytic gradient: dict(loss (S,), grad (S,A,T,2)[, sd (S,A,T)]) device tensors."""
        b = self.backend
        idx = {"repeller": 0, "no_offroad": 1}[which]
        if radius is None:
            radius = 6.0 if idx == 0 else 1.0
        x = b.to_device(xy, np.float32)
        T = int(x.shape[2])
        assert tuple(x.shape) == (self.S, self.A, T, 2)
        out = dict(loss=b.empty((self.S,), np.float32), grad=b.empty((self.S, self.A, T, 2), np.float32))
        if with_sd and idx == 1:
            out["sd"] = b.empty((self.S, self.A, T), np.float32)
        self._check(self.lib.lcsim_b200_guidance_cost(self._h, idx, b.ptr(x), T, float(radius), float(eps), b.ptr(out["loss"]),
                                                      b.ptr(out["grad"]), b.ptr(out["sd"]) if "sd" in out else None,
                                                      b.stream()), "guidance_cost")
        self._keep = x
        return out

    def plan_metrics(self, traj5, overlap_mask=None, offroad_mask=None):
        """compute_overlap_rate / compute_offroad_rate (motion_diff/metrics/overlap.py:7-43, roadgraph.py:36-88) of the
        plans traj5 (S,A,T,5) = x, y, length, width, yaw (float32), per scenario: dict(overlap_rate, overlap_count,
        offroad_rate, offroad_count) device tensors.  The masks are (S,A); offroad_mask should carry the vehicle mask."""
        b = self.backend
        x = b.to_device(traj5, np.float32)
        T = int(x.shape[2])
        assert tuple(x.shape) == (self.S, self.A, T, 5)
        m1 = None if overlap_mask is None else b.to_device(overlap_mask, np.uint8)
        m2 = None if offroad_mask is None else b.to_device(offroad_mask, np.uint8)
        out = dict(overlap_rate=b.empty((self.S,), np.float32), overlap_count=b.empty((self.S,), np.int32),
                   offroad_rate=b.empty((self.S,), np.float32), offroad_count=b.empty((self.S,), np.int32))
        self._check(self.lib.lcsim_b200_plan_metrics(
            self._h, b.ptr(x), T, None if m1 is None else b.ptr(m1), None if m2 is None else b.ptr(m2),
            b.ptr(out["overlap_rate"]), b.ptr(out["overlap_count"]), b.ptr(out["offroad_rate"]), b.ptr(out["offroad_count"]),
            b.stream()), "plan_metrics")
        self._keep = (x, m1, m2)
        return out

    # ------------------------------------------------------------------ aggregates
    def episode_stats(self):
        """(8,) int64 device tensor: this rank's sums (see native.STAT_KEYS)."""
        self._check(self.lib.lcsim_b200_episode_stats(self._h, self.backend.ptr(self._stats_out), self.backend.stream()),
                    "episode_stats")
        return self._stats_out

    def ring_ordered(self):
        """History ring in the reference's shift-register order (oldest first, newest last), (S,A,11,5)."""
        H = native.H_STEPS
        xp = self.backend.torch if hasattr(self.backend, "torch") else None
        if xp is not None:
            idx = (self.ring_head.long().unsqueeze(-1) + xp.arange(H, device=self.ring.device)) % H
            return xp.gather(self.ring, 2, idx.unsqueeze(-1).expand(-1, -1, -1, 5))
        idx = (self.ring_head.astype(np.int64)[..., None] + np.arange(H)) % H
        return np.take_along_axis(self.ring, idx[..., None], axis=2)
