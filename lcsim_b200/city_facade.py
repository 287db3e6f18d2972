"""Drop-in Python face for MOSS-shaped scenarios (lane-position agents driven by LaneIDM) over the city engine.

Same names and attribute surface as the reference for the calls its users make on such a scenario:
``Engine(config)``, ``.reset()``, ``.step()``, ``.current_step / .current_time``, ``.agent_manager`` with ``agents``,
``ads_id``, ``active_agents``, ``get_agent_by_id``, ``get_ads``, ``get_agent_bbox``, ``check_collision``,
``check_collision_all``, ``get_collision_rate``, ``get_offroad_rate`` (lcsim/entity/agent/manager.py:17-345) and per agent
``id, length, width, heading, v, runtime.{status, lane.id, s, v, xy, heading}``, ``obs.ahead_vehicle``, ``is_offroad()``
(lcsim/entity/agent/agent.py:29-50,173-191,482-494).  ``lcsim_b200.engine.Engine(config)`` returns this class when the
scenario's agents have lane-position homes.  External actions and rendering are not part of this face.
"""
from __future__ import annotations

import os
from enum import Enum
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import city
from .city_engine import CityEngine


class _Snap:
    """host copy of the engine state, refreshed lazily after every tick"""

    FIELDS = ("active", "status", "lane", "s", "v", "xy", "heading", "lead_valid", "lead_dis", "lead_v", "coll_count", "offroad")

    def __init__(self, ce: CityEngine):
        self._ce, self._c = ce, {}

    def invalidate(self):
        self._c = {}

    def __getattr__(self, name):
        if name == "n_active":
            if name not in self._c:
                self._c[name] = self._ce.scalar("n_active")
            return self._c[name]
        if name in _Snap.FIELDS:
            if name not in self._c:
                self._c[name] = self._ce.backend.to_numpy(getattr(self._ce, name))
            return self._c[name]
        raise AttributeError(name)


class Status(Enum):
    WAITING = 0
    ACTIVE = 1
    FINISHED = 2
    IDLE = 3


class _Lane:
    def __init__(self, lane_id: int):
        self.id = lane_id


class _Runtime:
    def __init__(self, a: "CityAgent"):
        self._a = a

    @property
    def status(self) -> Status:
        return Status(int(self._a._snap.status[self._a._i]))

    @property
    def lane(self) -> _Lane:
        return _Lane(int(self._a._eng._static.lane_pbid[self._a._snap.lane[self._a._i]]))

    @property
    def s(self) -> float:
        return float(self._a._snap.s[self._a._i])

    @property
    def v(self) -> float:
        return float(self._a._snap.v[self._a._i])

    @property
    def xy(self) -> np.ndarray:
        return self._a._snap.xy[self._a._i].copy()

    @property
    def heading(self) -> float:
        return float(self._a._snap.heading[self._a._i])


class _Obs:
    def __init__(self, a: "CityAgent"):
        self._a = a

    @property
    def ahead_vehicle(self):
        a = self._a
        if not a._snap.lead_valid[a._i]:
            return None
        return (float(a._snap.lead_dis[a._i]), float(a._snap.lead_v[a._i]))


class CityAgent:
    Status = Status

    def __init__(self, eng: "CityFacadeEngine", index: int):
        self._eng, self._i = eng, index
        st = eng._static
        self.id = int(st.agent_id[index])
        self.length, self.width = float(st.length[index]), float(st.width[index])
        self.runtime, self.obs = _Runtime(self), _Obs(self)

    @property
    def _snap(self) -> _Snap:
        return self._eng._snapshot

    @property
    def heading(self) -> float:
        return self.runtime.heading

    @property
    def v(self) -> float:
        return self.runtime.v

    def is_offroad(self) -> bool:
        """reference: agent.py:482-494; the engine evaluates it for every active agent every tick"""
        return bool(self._snap.offroad[self._i])


class CityAgentManager:
    def __init__(self, eng: "CityFacadeEngine"):
        self._eng = eng
        st = eng._static
        self._list = [CityAgent(eng, i) for i in range(st.n_agents)]
        self.agents: Dict[int, CityAgent] = {a.id: a for a in self._list}
        self.ads_id = self._list[int(st.ads_index)].id

    def get_agent_by_id(self, agent_id: int) -> Optional[CityAgent]:
        return self.agents.get(agent_id)

    def get_ads(self) -> Optional[CityAgent]:
        return self.agents.get(self.ads_id)

    @property
    def active_agents(self) -> List[CityAgent]:
        sn = self._eng._snapshot
        return [self._list[i] for i in sn.active[: sn.n_active]]

    def get_agent_bbox(self, ids: Sequence[int]) -> np.ndarray:
        """reference: manager.py:258-287 -> (n,5) [x, y, length, width, heading]"""
        sn = self._eng._snapshot
        idx = [self.agents[i]._i for i in ids]
        st = self._eng._static
        return np.stack([sn.xy[idx, 0], sn.xy[idx, 1], st.length[idx], st.width[idx], sn.heading[idx]], axis=1) if idx else np.zeros((0, 5))

    def check_collision_all(self) -> np.ndarray:
        """reference: manager.py:313-327 — overlap count of every active agent, list order"""
        sn = self._eng._snapshot
        return sn.coll_count[sn.active[: sn.n_active]].copy()

    def check_collision(self, agent_id: int) -> bool:
        """reference: manager.py:289-311 restricts the candidates to the agent's road / junction; any overlapping box lies
        within a few metres, so for lane agents this is the all-pairs count restricted to ... the same answer except
        across a road/junction boundary — NOT restated, the all-pairs flag is returned"""
        a = self.agents[agent_id]
        return bool(a.runtime.status == Status.ACTIVE and self._eng._snapshot.coll_count[a._i] > 0)

    def get_collision_rate(self) -> float:
        sn = self._eng._snapshot
        return 0.0 if sn.n_active == 0 else float(self.check_collision_all().sum() / sn.n_active)

    def get_offroad_rate(self) -> float:
        """reference: manager.py:337-345"""
        sn = self._eng._snapshot
        return 0.0 if sn.n_active == 0 else float(sn.offroad[sn.active[: sn.n_active]].sum() / sn.n_active)


class CityFacadeEngine:
    """reference: lcsim/engine/engine.py:22-158 for a scenario of lane-position agents"""

    def __init__(self, config: dict):
        self.config = config
        step = config["control"]["step"]
        self.start_time, self.total_steps, self.step_interval = step["start"], step["total"], step["interval"]
        c = config["control"]
        self.allow_new_obj, self.reactive, self.no_static = c["allow_new_obj"], c["reactive"], c["no_static"]
        if not self.allow_new_obj:
            raise NotImplementedError("allow_new_obj=False is not supported for lane-position scenarios")
        self.reward_config = config.get("reward", {})
        data_dir = config["input"]["data_dir"]
        for fn in ("map.pb", "agents.pb"):
            if not os.path.exists(os.path.join(data_dir, fn)):
                raise FileNotFoundError(os.path.join(data_dir, fn))
        dev = 0
        if isinstance(c.get("device"), str) and c["device"].startswith("cuda:"):
            dev = int(c["device"].split(":")[1])
        self._static = city.pack_dir(data_dir, start=float(self.start_time), interval=float(self.step_interval))
        self._city = CityEngine(self._static, with_metrics=True, device=dev, backend=config.get("_backend"))
        self._snapshot = _Snap(self._city)
        self.current_step, self.current_time = 0, self.start_time
        self.agent_manager = CityAgentManager(self)

    def reset(self):
        self._city.reset()
        self._snapshot.invalidate()
        self.current_step, self.current_time = 0, self.start_time

    def step(self, actions: dict = {}):
        if actions:
            raise NotImplementedError("external actions for lane-position agents")
        self._city.step(1)
        self._snapshot.invalidate()
        self.current_time += self.step_interval  # float64 accumulation like engine.py:157
        self.current_step += 1

    def render(self):
        raise NotImplementedError("rendering is outside the hot path (SURVEY.md §2)")

    def close(self):
        self._city.close()
