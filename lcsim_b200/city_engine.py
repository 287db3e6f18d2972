"""City engine: one MOSS-shaped scenario whose agents the reference drives with LaneIDM.

Host-side owner of a ``lcsim_b200_city`` handle (include/lcsim_b200.h).  Mirrors the reference's ``Engine.reset`` /
``Engine.step`` (lcsim/engine/engine.py:132-158) for a scenario of lane-position agents: ``AgentManager.update`` /
``check_departure_and_arrival`` / ``prepare_obs`` (lcsim/entity/agent/manager.py:50-115), ``AgentManager.check_collision_all``
(:313-327) and ``Agent.is_offroad`` (lcsim/entity/agent/agent.py:482-494) for every active agent, every tick.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import native
from .city import CityStatic

SCALARS = {"n_active": 0, "wait_ptr": 1, "tick": 2, "lc_required": 6, "lc_error": 7}


class CityEngine:
    def __init__(self, static: CityStatic, with_metrics: bool = True, device: int = 0, backend=None):
        self.backend = backend if backend is not None else native.TorchCudaBackend(device)
        self.lib = self.backend.lib
        self.static = static
        self.N, self.n_lanes = int(static.n_agents), int(static.n_lanes)
        st = native.CityStaticC()
        st.start, st.interval = float(static.start), float(static.interval)
        st.n_lanes, st.n_agents = self.n_lanes, self.N
        st.n_esets, st.n_groups = len(static.eset_start) - 1, len(static.grp_start) - 1
        keep = []
        for name, dt in native.CITY_STATIC_FIELDS:
            arr = np.ascontiguousarray(getattr(static, name), dtype=dt)
            keep.append(arr)
            setattr(st, name, arr.ctypes.data)
        self._h = C.c_void_p()
        rc = self.lib.lcsim_b200_city_create(C.byref(st), int(getattr(self.backend, "device_index", 0)), int(with_metrics),
                                             C.byref(self._h))
        if rc != 0:
            msg = self.lib.lcsim_b200_city_last_error(self._h).decode() if self._h else "create failed"
            self.close()
            raise native.NativeLibraryError(f"lcsim_b200_city_create: {rc}: {msg}")
        del keep
        v = native.CityViews()
        self._check(self.lib.lcsim_b200_city_views(self._h, C.byref(v)), "views")
        N, H = self.N, native.H_STEPS
        shapes = {"scalars": (16,), "is_lc": (N,), "active": (N,), "status": (N,), "lane": (N,), "s": (N,), "v": (N,), "xy": (N, 2),
                  "heading": (N,), "lead_valid": (N,), "lead_dis": (N,), "lead_v": (N,), "coll_count": (N,),
                  "offroad": (N,), "ring": (N, H, 5), "ring_head": (N,), "ring_count": (N,), "stats": (native.N_STATS,)}
        for name, typestr in native.CITY_VIEW_FIELDS:
            setattr(self, name, self.backend.view(getattr(v, name), shapes[name], typestr))

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise native.NativeLibraryError(f"lcsim_b200_city_{what}: {rc}: {self.lib.lcsim_b200_city_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.lcsim_b200_city_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._check(self.lib.lcsim_b200_city_reset(self._h, self.backend.stream()), "reset")

    def step(self, n_ticks: int = 1):
        self._check(self.lib.lcsim_b200_city_step(self._h, int(n_ticks), self.backend.stream()), "step")

    @property
    def launches(self) -> int:
        return int(self.lib.lcsim_b200_city_launch_count(self._h))

    def scalar(self, name: str) -> int:
        return int(self.backend.to_numpy(self.scalars)[SCALARS[name]])

    def ring_ordered(self):
        H = native.H_STEPS
        ring, head = self.backend.to_numpy(self.ring), self.backend.to_numpy(self.ring_head)
        idx = (head.astype(np.int64)[:, None] + np.arange(H)) % H
        return np.take_along_axis(ring, idx[..., None], axis=1)
