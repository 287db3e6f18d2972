"""ctypes binding of the CPU oracle (oracle/lcsim_oracle.c).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module; lcsim_b200/ never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblcsim_oracle.so")

ACT_NONE, ACT_BICYCLE, ACT_DELTA, ACT_STATE = 0, 1, 2, 3
WAITING, ACTIVE, FINISHED, IDLE = 0, 1, 2, 3
H, ROUTE = 11, 10


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lcsim_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB_PATH


class _Static(C.Structure):
    _fields_ = [
        ("S", C.c_int32), ("A", C.c_int32), ("T", C.c_int32), ("E", C.c_int32),
        ("start", C.c_double), ("interval", C.c_double),
        ("total_steps", C.c_int32),
        ("reactive", C.c_int32), ("no_static", C.c_int32), ("allow_new_obj", C.c_int32),
        ("n_agents", C.c_void_p), ("ads_index", C.c_void_p), ("agent_id", C.c_void_p), ("agent_type", C.c_void_p),
        ("length", C.c_void_p), ("width", C.c_void_p),
        ("dynamic", C.c_void_p),
        ("traj_len", C.c_void_p), ("depart_tick", C.c_void_p), ("wait_order", C.c_void_p),
        ("home_xy", C.c_void_p), ("traj_xy", C.c_void_p), ("traj_vel", C.c_void_p), ("traj_h", C.c_void_p),
        ("n_edge", C.c_void_p),
        ("edge_xy", C.c_void_p), ("edge_dir", C.c_void_p),
    ]


class _Views(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "tick", "n_active", "active", "status", "cursor", "ring_valid", "coll_count", "nearest_edge", "traj_len",
        "x", "y", "h", "v", "lead_dis", "lead_v", "ring", "traj_xy", "traj_vel", "traj_h",
        "lead_valid", "offroad", "agent_steps")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_create.argtypes = [C.POINTER(_Static)]
        _lib.orc_destroy.argtypes = [C.c_void_p]
        _lib.orc_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib.orc_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.orc_metrics.argtypes = [C.c_void_p]
        _lib.orc_views.argtypes = [C.c_void_p, C.POINTER(_Views)]
        _lib.orc_route.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_rewards.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib.orc_is_offroad.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_is_offroad.restype = C.c_int
        _lib.orc_check_collision.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _lib.orc_check_collision.restype = C.c_int
        _lib.orc_set_ref_trajectory.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_rollout_mt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib.orc_rollout_mt.restype = C.c_int64
        _lib.orc_test_np_sum.argtypes = [C.c_void_p, C.c_int64]
        _lib.orc_test_np_sum.restype = C.c_double
        _lib.orc_test_wrap_angle.argtypes = [C.c_double]
        _lib.orc_test_wrap_angle.restype = C.c_double
        _lib.orc_s_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """CPU oracle engine over a packed StaticBatch (lcsim_b200.packer.StaticBatch)."""

    def __init__(self, static, reactive: bool = False, no_static: bool = False, allow_new_obj: bool = True):
        self.static = static  # keeps the numpy arrays alive
        st = _Static()
        st.S, st.A, st.T, st.E = static.S, static.A, static.T, static.E
        st.start, st.interval, st.total_steps = static.start, static.interval, static.total_steps
        st.reactive, st.no_static, st.allow_new_obj = int(reactive), int(no_static), int(allow_new_obj)
        for name, dt in (("n_agents", np.int32), ("ads_index", np.int32), ("agent_id", np.int32),
                         ("agent_type", np.int32), ("length", np.float64), ("width", np.float64),
                         ("dynamic", np.uint8), ("traj_len", np.int32), ("depart_tick", np.int32),
                         ("wait_order", np.int32), ("home_xy", np.float64), ("traj_xy", np.float64),
                         ("traj_vel", np.float64), ("traj_h", np.float64), ("n_edge", np.int32),
                         ("edge_xy", np.float64), ("edge_dir", np.float64)):
            arr = getattr(static, name)
            assert arr.dtype == dt and arr.flags["C_CONTIGUOUS"], name
            setattr(st, name, arr.ctypes.data)
        self._st = st
        self._h = lib().orc_create(C.byref(st))
        v = _Views()
        lib().orc_views(self._h, C.byref(v))
        S, A, T = static.S, static.A, static.T

        def view(ptr, shape, dtype):
            n = int(np.prod(shape))
            buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype).reshape(shape)

        self.tick = view(v.tick, (S,), np.int32)
        self.n_active = view(v.n_active, (S,), np.int32)
        self.active = view(v.active, (S, A), np.int32)
        self.status = view(v.status, (S, A), np.int32)
        self.cursor = view(v.cursor, (S, A), np.int32)
        self.ring_valid = view(v.ring_valid, (S, A), np.int32)
        self.coll_count = view(v.coll_count, (S, A), np.int32)
        self.nearest_edge = view(v.nearest_edge, (S, A), np.int32)
        self.traj_len = view(v.traj_len, (S, A), np.int32)
        self.x = view(v.x, (S, A), np.float64)
        self.y = view(v.y, (S, A), np.float64)
        self.h = view(v.h, (S, A), np.float64)
        self.v = view(v.v, (S, A), np.float64)
        self.lead_dis = view(v.lead_dis, (S, A), np.float64)
        self.lead_v = view(v.lead_v, (S, A), np.float64)
        self.lead_valid = view(v.lead_valid, (S, A), np.uint8)
        self.ring = view(v.ring, (S, A, H, 5), np.float64)
        self.offroad = view(v.offroad, (S, A), np.uint8)
        self.agent_steps = view(v.agent_steps, (S,), np.int64)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def reset(self, scenario_ids=None):
        if scenario_ids is None:
            lib().orc_reset(self._h, None, 0)
        else:
            ids = np.ascontiguousarray(scenario_ids, np.int32)
            lib().orc_reset(self._h, _ptr(ids), len(ids))

    def step(self, act_agent=None, act_type=None, act_data=None):
        """act_agent/act_type: (S,K) int32, act_data: (S,K,5) float64; None -> own policies."""
        if act_agent is None:
            lib().orc_step(self._h, None, None, None, 0)
            return
        aa = np.ascontiguousarray(act_agent, np.int32)
        at = np.ascontiguousarray(act_type, np.int32)
        ad = np.ascontiguousarray(act_data, np.float64)
        K = aa.shape[1]
        assert aa.shape == at.shape == (self.static.S, K) and ad.shape == (self.static.S, K, 5)
        lib().orc_step(self._h, _ptr(aa), _ptr(at), _ptr(ad), K)

    def metrics(self):
        lib().orc_metrics(self._h)

    def route(self, s: int, a: int) -> np.ndarray:
        out = np.zeros((ROUTE, 2))
        lib().orc_route(self._h, s, a, _ptr(out))
        return out

    def rewards(self, s: int, a: int):
        terms = np.zeros(6)
        term = np.zeros(1, np.int32)
        lib().orc_rewards(self._h, s, a, _ptr(terms), _ptr(term))
        return terms, bool(term[0])

    def is_offroad(self, s: int, a: int):
        ne = np.zeros(1, np.int32)
        r = lib().orc_is_offroad(self._h, s, a, _ptr(ne))
        return bool(r), int(ne[0])

    def check_collision(self, s: int, a: int) -> bool:
        return bool(lib().orc_check_collision(self._h, s, a))

    def set_ref_trajectory(self, s: int, a: int, xy, vel, heading):
        xy = np.ascontiguousarray(xy, np.float32)
        vel = np.ascontiguousarray(vel, np.float32)
        heading = np.ascontiguousarray(heading, np.float32)
        lib().orc_set_ref_trajectory(self._h, s, a, len(heading), _ptr(xy), _ptr(vel), _ptr(heading))

    def rollout_mt(self, n_ticks: int, with_metrics: bool, n_threads: int) -> int:
        return int(lib().orc_rollout_mt(self._h, n_ticks, int(with_metrics), n_threads))


def np_sum(a: np.ndarray) -> float:
    a = np.ascontiguousarray(a, np.float64)
    return float(lib().orc_test_np_sum(_ptr(a), a.size))


def wrap_angle(a: float) -> float:
    return float(lib().orc_test_wrap_angle(a))


def s_table(xy: np.ndarray, f32: bool = False) -> np.ndarray:
    xy = np.ascontiguousarray(xy, np.float64)
    out = np.zeros(len(xy))
    lib().orc_s_table(_ptr(xy), len(xy), _ptr(out), int(f32))
    return out
