"""ctypes binding of the C ABI in ``include/lcsim_b200.h`` (``lcsim_b200/csrc/lcsim_b200.cu``).

The product path is CUDA only: :func:`load` returns the sm_100a library or raises
:class:`NativeLibraryError`; nothing in this package computes on the CPU when the library
(or a GPU) is missing.  PyTorch appears only as the owner of streams and as a zero-copy
view onto the engine's device buffers (``__cuda_array_interface__``).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
SRC = os.path.join(_HERE, "csrc", "lcsim_b200.cu")
HDRS = [os.path.join(_HERE, "csrc", "lcsim_core.cuh"), os.path.join(_HERE, "csrc", "lcsim_city.cuh"),
        os.path.join(_HERE, "csrc", "lcsim_plan.cuh"), os.path.join(_HERE, "csrc", "lcsim_loader.hpp"),
        os.path.join(_HERE, "csrc", "lcsim_encoder.cuh"),
        os.path.join(ROOT, "include", "lcsim_b200.h")]
LIB_PATH = os.path.join(_HERE, "liblcsim_b200.so")

ABI_VERSION = 2
H_STEPS, ROUTE_LEN, N_REWARD_TERMS, N_STATS = 11, 10, 6, 8
WAITING, ACTIVE, FINISHED, IDLE = 0, 1, 2, 3
ACT_NONE, ACT_BICYCLE, ACT_DELTA, ACT_STATE = 0, 1, 2, 3
REWARD_KEYS = ("forward", "route_progress", "smooth", "collision", "offroad", "destination")
STAT_KEYS = ("agent_steps", "active_sum", "collision_sum", "offroad_sum", "arrivals", "departures", "ticks", "reserved")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-Xcompiler", "-ffp-contract=off"]


class NativeLibraryError(RuntimeError):
    pass


def _nvcc() -> str | None:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [SRC] + HDRS
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    nvcc = _nvcc()
    if nvcc is None:
        raise NativeLibraryError("nvcc not found: cannot build lcsim_b200/liblcsim_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [SRC, "-o", LIB_PATH + ".tmp"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise NativeLibraryError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


class BatchDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("S", C.c_int32), ("A", C.c_int32), ("T", C.c_int32), ("E", C.c_int32),
                ("start", C.c_double), ("interval", C.c_double), ("total_steps", C.c_int32),
                ("reactive", C.c_int32), ("no_static", C.c_int32), ("allow_new_obj", C.c_int32),
                ("device", C.c_int32), ("with_metrics", C.c_int32), ("grid_cell", C.c_float)]


STATIC_FIELDS = (("n_agents", np.int32), ("ads_index", np.int32), ("agent_type", np.int32), ("length", np.float64),
                 ("width", np.float64), ("dynamic", np.uint8), ("traj_len", np.int32), ("depart_tick", np.int32),
                 ("wait_order", np.int32), ("home_xy", np.float64), ("traj_xy", np.float64), ("traj_vel", np.float64),
                 ("traj_h", np.float64), ("n_edge", np.int32), ("edge_xy", np.float64), ("edge_dir", np.float64),
                 ("origin", np.float64), ("edge_poly_id", np.int32))


class StaticHost(C.Structure):
    _fields_ = [(n, C.c_void_p) for n, _ in STATIC_FIELDS]


VIEW_FIELDS = (("tick", "<i4"), ("n_active", "<i4"), ("active", "<i4"), ("status", "<i4"), ("cursor", "<i4"),
               ("ring_count", "<i4"), ("ring_head", "<i4"), ("pose64", "<f8"), ("pose32", "<f4"), ("ring", "<f8"),
               ("lead_valid", "|u1"), ("lead_dis", "<f8"), ("lead_v", "<f8"), ("coll_count", "<i4"),
               ("nearest_edge", "<i4"), ("offroad", "|u1"), ("traj_len", "<i4"), ("traj_xy", "<f8"),
               ("traj_vel", "<f8"), ("traj_h", "<f8"), ("stats", "<i8"), ("origin", "<f8"))


class StateViews(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("S", "A", "A_stride", "T", "E", "H")] + [(n, C.c_void_p) for n, _ in VIEW_FIELDS]


ENV_OUT_DTYPE = np.dtype([("reward", "<f8"), ("terms", "<f8", (N_REWARD_TERMS,)), ("route", "<f8", (ROUTE_LEN, 2)),
                          ("terminated", "u1"), ("truncated", "u1"), ("finished", "u1"), ("pad", "u1", (5,))])
assert ENV_OUT_DTYPE.itemsize == 224


class ObsDesc(C.Structure):
    _fields_ = [("ego", C.c_void_p), ("radius", C.c_double), ("k_agents", C.c_int32), ("k_polys", C.c_int32),
                ("sort_by_distance", C.c_int32), ("nbr_agent", C.c_void_p), ("nbr_agent_feat", C.c_void_p),
                ("n_nbr_agent", C.c_void_p), ("nbr_poly", C.c_void_p), ("nbr_poly_feat", C.c_void_p),
                ("n_nbr_poly", C.c_void_p), ("history", C.c_void_p), ("history_agent", C.c_void_p)]


class EncoderDesc(C.Structure):
    _fields_ = [("a2a_radius", C.c_double), ("pl2a_radius", C.c_double), ("time_span", C.c_int32), ("cap_a2a", C.c_int32),
                ("cap_pl2a", C.c_int32)] + [(n, C.c_void_p) for n in (
                    "agent", "x_a", "r_t", "r_t_valid", "a2a_offset", "a2a_src", "a2a_feat", "n_a2a", "pl2a_offset", "pl2a_agent",
                    "pl2a_feat", "n_pl2a")]


class RolloutOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("tick_log", "pose32", "status", "coll_count", "nearest_edge", "offroad", "stats")]


class LaneDesc(C.Structure):
    _fields_ = [("top_k", C.c_int32), ("lane_pt", C.c_void_p), ("lane_id", C.c_void_p), ("lane_dist", C.c_void_p),
                ("onlane_dist", C.c_void_p), ("heading_diff", C.c_void_p)]


CITY_STATIC_FIELDS = (
    ("lane_length", np.float64), ("lane_max_speed", np.float64), ("lane_road", np.int32), ("lane_junction", np.int32),
    ("lane_offset", np.int32), ("lane_succ0", np.int32), ("lane_left", np.int32), ("lane_right", np.int32),
    ("lane_width", np.float64), ("lane_node_start", np.int32), ("node_xy", np.float64),
    ("node_cum", np.float64), ("node_dir", np.float64), ("lane_eset", np.int32), ("eset_start", np.int32),
    ("edge_xy", np.float64), ("edge_dir", np.float64), ("grp_start", np.int32), ("grp_lane", np.int32),
    ("grp_pre_offset", np.int32), ("length", np.float64), ("width", np.float64), ("idm", np.float64),
    ("home_lane", np.int32), ("home_s", np.float64), ("end_road", np.int32), ("end_offset", np.int32),
    ("end_s", np.float64), ("depart_tick", np.int32), ("wait_order", np.int32), ("route_start", np.int32),
    ("route_grp", np.int32))


class CityStaticC(C.Structure):
    _fields_ = ([("start", C.c_double), ("interval", C.c_double), ("n_lanes", C.c_int32), ("n_agents", C.c_int32),
                 ("n_esets", C.c_int32), ("n_groups", C.c_int32)] + [(n, C.c_void_p) for n, _ in CITY_STATIC_FIELDS])


CITY_VIEW_FIELDS = (("scalars", "<i4"), ("is_lc", "<i4"), ("active", "<i4"), ("status", "<i4"), ("lane", "<i4"), ("s", "<f8"), ("v", "<f8"),
                    ("xy", "<f8"), ("heading", "<f8"), ("lead_valid", "|u1"), ("lead_dis", "<f8"), ("lead_v", "<f8"),
                    ("coll_count", "<i4"), ("offroad", "|u1"), ("ring", "<f8"), ("ring_head", "<i4"),
                    ("ring_count", "<i4"), ("stats", "<i8"))


class CityViews(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("N", "n_lanes", "H")] + [(n, C.c_void_p) for n, _ in CITY_VIEW_FIELDS]


EXPORTS = ("lcsim_b200_abi_version", "lcsim_b200_create", "lcsim_b200_upload_static", "lcsim_b200_reset",
           "lcsim_b200_step", "lcsim_b200_step_masked", "lcsim_b200_rollout", "lcsim_b200_rollout_log", "lcsim_b200_rollout_host", "lcsim_b200_set_ref_trajectory", "lcsim_b200_rewards",
           "lcsim_b200_env_step", "lcsim_b200_env_step_device", "lcsim_b200_set_env_obs", "lcsim_b200_upload_polylines", "lcsim_b200_build_obs", "lcsim_b200_build_encoder_inputs",
           "lcsim_b200_upload_centrelines", "lcsim_b200_project_lanes", "lcsim_b200_guidance_cost",
           "lcsim_b200_plan_metrics", "lcsim_b200_load_dirs", "lcsim_b200_scenarios_dims", "lcsim_b200_scenarios_static",
           "lcsim_b200_scenarios_agent_id", "lcsim_b200_scenarios_junction_id", "lcsim_b200_scenarios_polylines",
           "lcsim_b200_scenarios_error", "lcsim_b200_scenarios_free", "lcsim_b200_views", "lcsim_b200_episode_stats",
           "lcsim_b200_launch_count", "lcsim_b200_last_error", "lcsim_b200_destroy",
           "lcsim_b200_city_create", "lcsim_b200_city_reset", "lcsim_b200_city_step", "lcsim_b200_city_views",
           "lcsim_b200_city_launch_count", "lcsim_b200_city_last_error", "lcsim_b200_city_destroy")


def bind(path: str) -> C.CDLL:
    """dlopen + prototypes.  Raises if a symbol declared in include/lcsim_b200.h is missing."""
    lib = C.CDLL(path)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise NativeLibraryError(f"{path}: missing export {name}")
    vp, i32 = C.c_void_p, C.c_int
    lib.lcsim_b200_abi_version.restype = i32
    lib.lcsim_b200_create.argtypes = [C.POINTER(BatchDesc), C.POINTER(vp)]
    lib.lcsim_b200_upload_static.argtypes = [vp, C.POINTER(StaticHost)]
    lib.lcsim_b200_reset.argtypes = [vp, vp, vp]
    lib.lcsim_b200_step.argtypes = [vp, vp, vp, vp, i32, vp]
    lib.lcsim_b200_step_masked.argtypes = [vp, vp, vp, vp, vp, i32, vp]
    lib.lcsim_b200_rollout.argtypes = [vp, i32, vp]
    lib.lcsim_b200_rollout_log.argtypes = [vp, i32, vp, vp]
    lib.lcsim_b200_rollout_host.argtypes = [vp, vp, i32, C.POINTER(RolloutOut), i32, vp]
    lib.lcsim_b200_set_ref_trajectory.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp]
    lib.lcsim_b200_rewards.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.lcsim_b200_env_step.argtypes = [vp, vp, vp, vp, i32, vp]
    lib.lcsim_b200_set_env_obs.argtypes = [vp, C.POINTER(ObsDesc)]
    lib.lcsim_b200_env_step_device.argtypes = [vp, vp, vp, vp, vp]
    lib.lcsim_b200_upload_polylines.argtypes = [vp, vp, vp, i32]
    lib.lcsim_b200_build_obs.argtypes = [vp, C.POINTER(ObsDesc), vp]
    lib.lcsim_b200_build_encoder_inputs.argtypes = [vp, C.POINTER(EncoderDesc), vp]
    lib.lcsim_b200_upload_centrelines.argtypes = [vp, vp, vp, vp, i32]
    lib.lcsim_b200_project_lanes.argtypes = [vp, C.POINTER(LaneDesc), vp]
    lib.lcsim_b200_guidance_cost.argtypes = [vp, i32, vp, i32, C.c_float, C.c_float, vp, vp, vp, vp]
    lib.lcsim_b200_plan_metrics.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.lcsim_b200_load_dirs.argtypes = [C.POINTER(C.c_char_p), i32, C.c_double, C.c_double, C.c_int32, i32, C.POINTER(vp)]
    lib.lcsim_b200_scenarios_dims.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.lcsim_b200_scenarios_static.argtypes = [vp, C.POINTER(StaticHost)]
    lib.lcsim_b200_scenarios_agent_id.argtypes = [vp]
    lib.lcsim_b200_scenarios_agent_id.restype = vp
    lib.lcsim_b200_scenarios_junction_id.argtypes = [vp]
    lib.lcsim_b200_scenarios_junction_id.restype = vp
    lib.lcsim_b200_scenarios_polylines.argtypes = [vp, C.POINTER(vp)]
    lib.lcsim_b200_scenarios_polylines.restype = vp
    lib.lcsim_b200_scenarios_error.argtypes = [vp]
    lib.lcsim_b200_scenarios_error.restype = C.c_char_p
    lib.lcsim_b200_scenarios_free.argtypes = [vp]
    lib.lcsim_b200_scenarios_free.restype = None
    lib.lcsim_b200_views.argtypes = [vp, C.POINTER(StateViews)]
    lib.lcsim_b200_episode_stats.argtypes = [vp, vp, vp]
    lib.lcsim_b200_launch_count.argtypes = [vp]
    lib.lcsim_b200_launch_count.restype = C.c_int64
    lib.lcsim_b200_last_error.argtypes = [vp]
    lib.lcsim_b200_last_error.restype = C.c_char_p
    lib.lcsim_b200_destroy.argtypes = [vp]
    lib.lcsim_b200_destroy.restype = None
    lib.lcsim_b200_city_create.argtypes = [C.POINTER(CityStaticC), i32, i32, C.POINTER(vp)]
    lib.lcsim_b200_city_reset.argtypes = [vp, vp]
    lib.lcsim_b200_city_step.argtypes = [vp, i32, vp]
    lib.lcsim_b200_city_views.argtypes = [vp, C.POINTER(CityViews)]
    lib.lcsim_b200_city_launch_count.argtypes = [vp]
    lib.lcsim_b200_city_launch_count.restype = C.c_int64
    lib.lcsim_b200_city_last_error.argtypes = [vp]
    lib.lcsim_b200_city_last_error.restype = C.c_char_p
    lib.lcsim_b200_city_destroy.argtypes = [vp]
    lib.lcsim_b200_city_destroy.restype = None
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("lcsim_b200_abi_version",):
            fn.restype = i32
    if lib.lcsim_b200_abi_version() != ABI_VERSION:
        raise NativeLibraryError(f"{path}: ABI version {lib.lcsim_b200_abi_version()} != {ABI_VERSION}")
    return lib


def load_dirs(lib: C.CDLL, dirs, start: float = 0.0, interval: float = 0.1, total_steps: int = 91, n_threads: int = 0):
    """lcsim_b200_load_dirs: scenario directories -> (packer.StaticBatch, [polyline centres per scenario], junction ids).
    The arrays are copies; the native object is freed before returning."""
    from .packer import StaticBatch

    arr = (C.c_char_p * len(dirs))(*[os.fsencode(d) for d in dirs])
    h = C.c_void_p()
    rc = lib.lcsim_b200_load_dirs(arr, len(dirs), float(start), float(interval), int(total_steps), int(n_threads), C.byref(h))
    try:
        if rc != 0:
            raise NativeLibraryError(f"lcsim_b200_load_dirs: {rc}: {lib.lcsim_b200_scenarios_error(h).decode()}")
        dims = (C.c_int32 * 5)()
        lib.lcsim_b200_scenarios_dims(h, dims)
        S, A, T, E, P = (int(x) for x in dims)
        sh = StaticHost()
        lib.lcsim_b200_scenarios_static(h, C.byref(sh))
        shapes = {"n_agents": (S,), "ads_index": (S,), "agent_type": (S, A), "length": (S, A), "width": (S, A), "dynamic": (S, A),
                  "traj_len": (S, A), "depart_tick": (S, A), "wait_order": (S, A), "home_xy": (S, A, 2), "traj_xy": (S, A, T, 2),
                  "traj_vel": (S, A, T, 2), "traj_h": (S, A, T), "n_edge": (S,), "edge_xy": (S, E, 2), "edge_dir": (S, E, 2),
                  "origin": (S, 2), "edge_poly_id": (S, E)}

        def take(ptr, shape, dt):
            n = int(np.prod(shape))
            buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(int(ptr))
            return np.frombuffer(buf, dtype=dt).reshape(shape).copy()

        kw = {name: take(getattr(sh, name), shapes[name], dt) for name, dt in STATIC_FIELDS}
        kw["agent_id"] = take(lib.lcsim_b200_scenarios_agent_id(h), (S, A), np.int32)
        n_poly_p = C.c_void_p()
        poly_p = lib.lcsim_b200_scenarios_polylines(h, C.byref(n_poly_p))
        n_poly = take(n_poly_p.value, (S,), np.int32)
        poly = take(poly_p, (S, P, 3), np.float64)
        junction = take(lib.lcsim_b200_scenarios_junction_id(h), (S,), np.int32)
        sb = StaticBatch(S=S, A=A, T=T, E=E, start=float(start), interval=float(interval), total_steps=int(total_steps), **kw)
        return sb, [poly[s, : n_poly[s]] for s in range(S)], junction
    finally:
        if h:
            lib.lcsim_b200_scenarios_free(h)


_LIB = None


def load() -> C.CDLL:
    """The CUDA library.  Built on first use when nvcc is present; otherwise it must already be in-tree."""
    global _LIB
    if _LIB is None:
        alt = os.environ.get("LCSIM_B200_LIB")  # experiments: another build of the same CUDA source
        if alt:
            _LIB = bind(alt)
            return _LIB
        if _nvcc() is not None:
            build()
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} is missing and nvcc is unavailable: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "lcsim_b200 has no CPU fallback.")
        _LIB = bind(LIB_PATH)
    return _LIB


class _CudaArray:
    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


class TorchCudaBackend:
    """Device memory and streams through PyTorch (plumbing only)."""

    name = "cuda"

    def __init__(self, device: int = 0):
        import torch

        if not torch.cuda.is_available():
            raise NativeLibraryError("no CUDA device visible: lcsim_b200 runs on the GPU only (no CPU fallback)")
        self.torch = torch
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        self.lib = load()

    def view(self, ptr: int, shape, typestr: str):
        return self.torch.as_tensor(_CudaArray(ptr, shape, typestr), device=self.device)

    def stream(self) -> int:
        return int(self.torch.cuda.current_stream(self.device).cuda_stream)

    def to_device(self, arr, dtype):
        t = self.torch
        if isinstance(arr, t.Tensor):
            return arr.to(device=self.device, dtype=getattr(t, np.dtype(dtype).name)).contiguous()
        return t.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).to(self.device)

    def empty(self, shape, dtype):
        return self.torch.empty(tuple(shape), dtype=getattr(self.torch, np.dtype(dtype).name), device=self.device)

    def ptr(self, x) -> int:
        return int(x.data_ptr())

    def to_numpy(self, x) -> np.ndarray:
        return x.detach().cpu().numpy()

    def sync(self):
        self.torch.cuda.synchronize(self.device)
