"""Drop-in facade with the reference's class and method names over the batched CUDA engine.

``Engine(config)`` takes the reference's config dict (lcsim/config/example.yml, read at
lcsim/engine/engine.py:59-86) and exposes ``reset / step / update / prepare / get_step_return`` plus an
``agent_manager`` whose ``Agent`` objects carry the attribute surface callers use
(``runtime.{status,xy,heading,v}``, ``history_states``, ``ref_trajectory``, ``is_offroad()``,
``get_route()``, ``set_ref_trajectory()``; lcsim/entity/agent/agent.py:86-104,173-191,462-494) and the
aggregates ``check_collision / check_collision_all / get_collision_rate / get_offroad_rate /
get_agent_bbox`` (lcsim/entity/agent/manager.py:258-345).  State lives on the GPU; attribute reads
materialise lazily from one device->host snapshot per tick.

``config["input"]["data_dir"]`` may also be a LIST of scenario directories: the engine then steps them as
one batch (one CUDA launch per tick for all of them) and ``engine.scenarios[i]`` is the per-scenario
facade; with a single directory ``engine`` itself is that facade, like the reference.

Not provided (out of the hot path): rendering, the MotionDiff encoder (``scene_emb`` is
zeros unless ``engine.scene_encoder`` is set to a callable), LaneIDM / lane-position agents.
"""
from __future__ import annotations

import os
from enum import Enum
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from . import mapview, native, packer, proto
from .batch import BatchEngine


class Action:
    """reference: lcsim/entity/agent/policy/type.py:74-165"""

    class ActionType(Enum):
        UNSPECIFIED = 0
        BICYCLE = 1
        DELTA = 2
        STATE = 3
        LANE = 4

    class BicycleAction:
        MAX_ACC = 6.0
        MAX_STEER = 0.3
        acc: float = 0.0
        steer: float = 0.0

        def __str__(self) -> str:
            return f"acc: {self.acc}, steer: {self.steer}"

        def init_with_normalized(self, acc: float, steer: float) -> None:
            self.acc = acc * self.MAX_ACC
            self.steer = steer * self.MAX_STEER

    class DeltaAction:
        dx: float = 0.0
        dy: float = 0.0
        dheading: float = 0.0

        def __str__(self) -> str:
            return f"dx: {self.dx}, dy: {self.dy}, dheading: {self.dheading}"

    class StateAction:
        x: float = 0.0
        y: float = 0.0
        vx: float = 0.0
        vy: float = 0.0
        heading: float = 0.0

        def __str__(self) -> str:
            return f"x: {self.x}, y: {self.y}, vx: {self.vx}, vy: {self.vy}, heading: {self.heading}"

    class LaneAction:
        acc: float = 0.0

    ACTION_MAP = {ActionType.BICYCLE: BicycleAction, ActionType.DELTA: DeltaAction,
                  ActionType.STATE: StateAction, ActionType.LANE: LaneAction}

    def __init__(self, type: "Action.ActionType", data):
        assert isinstance(data, self.ACTION_MAP[type])
        self.type = type
        self.data = data

    def _row(self) -> np.ndarray:
        d = self.data
        if self.type == Action.ActionType.BICYCLE:
            return np.array([d.acc, d.steer, 0.0, 0.0, 0.0])
        if self.type == Action.ActionType.DELTA:
            return np.array([d.dx, d.dy, d.dheading, 0.0, 0.0])
        if self.type == Action.ActionType.STATE:
            return np.array([d.x, d.y, d.vx, d.vy, d.heading])
        raise NotImplementedError("LANE actions belong to the LaneIDM path, which this engine does not cover")


class _Snapshot:
    """One device->host copy of the per-tick state of the whole batch, refreshed lazily after every tick."""

    FIELDS = ("n_active", "active", "status", "cursor", "pose64", "coll_count", "nearest_edge", "offroad",
              "lead_valid", "lead_dis", "lead_v", "ring_count", "traj_len")

    def __init__(self, be: BatchEngine):
        self.be = be
        self._d: dict = {}

    def invalidate(self):
        self._d = {}

    def __getattr__(self, name):
        if name.startswith("_") or name == "be":
            raise AttributeError(name)
        d = self.__dict__["_d"]
        if name not in d:
            be = self.__dict__["be"]
            if name == "ring":
                d[name] = be.backend.to_numpy(be.ring_ordered())
            else:
                d[name] = be.backend.to_numpy(getattr(be, name))
        return d[name]


class Agent:
    """reference: lcsim/entity/agent/agent.py (attribute surface only)"""

    class Status(Enum):
        WAITING = 0
        ACTIVE = 1
        FINISHED = 2
        IDLE = 3

    class Runtime:
        lane = None
        s = 0.0

        def __init__(self, agent: "Agent"):
            self._a = agent

        @property
        def status(self) -> "Agent.Status":
            return Agent.Status(int(self._a._snap.status[self._a._s, self._a._i]))

        @property
        def xy(self) -> np.ndarray:
            return self._a._snap.pose64[self._a._s, self._a._i, :2].copy()

        @property
        def heading(self) -> float:
            return float(self._a._snap.pose64[self._a._s, self._a._i, 2])

        @property
        def v(self) -> float:
            return float(self._a._snap.pose64[self._a._s, self._a._i, 3])

        def is_finished(self) -> bool:
            return self.status in (Agent.Status.IDLE, Agent.Status.FINISHED)

    class HistoryState:
        """(11,.) arrays in the reference's shift-register order (agent.py:52-84)."""

        def __init__(self, agent: "Agent"):
            self._a = agent

        def _ring(self):
            return self._a._snap.ring[self._a._s, self._a._i]

        @property
        def valid(self) -> np.ndarray:
            n = int(self._a._snap.ring_count[self._a._s, self._a._i])
            v = np.zeros((native.H_STEPS, 1), dtype=bool)
            v[native.H_STEPS - n:] = True
            return v

        @property
        def xyz(self) -> np.ndarray:
            r = self._ring()
            return np.concatenate([r[:, 0:2], np.zeros((native.H_STEPS, 1))], axis=1)

        @property
        def vel(self) -> np.ndarray:
            return self._ring()[:, 2:4].copy()

        @property
        def heading(self) -> np.ndarray:
            return self._ring()[:, 4:5].copy()

        @property
        def shape(self) -> np.ndarray:
            out = np.zeros((native.H_STEPS, 3))
            out[self.valid[:, 0], 0] = self._a.length
            out[self.valid[:, 0], 1] = self._a.width
            return out

        @property
        def type(self) -> np.ndarray:
            return np.ones(1, dtype=np.int32)  # hard-wired in the reference (agent.py:64)

    class _Trajectory:
        """reference: lcsim/entity/agent/policy/type.py:10-71 (read side)"""

        def __init__(self, agent: "Agent"):
            self._a = agent

        def _n(self):
            return int(self._a._snap.traj_len[self._a._s, self._a._i])

        @property
        def xy(self):
            a = self._a
            return a._eng._be.backend.to_numpy(a._eng._be.traj_xy[a._s, a._i, : self._n()])

        @property
        def vel(self):
            a = self._a
            return a._eng._be.backend.to_numpy(a._eng._be.traj_vel[a._s, a._i, : self._n()])

        @property
        def heading(self):
            a = self._a
            return a._eng._be.backend.to_numpy(a._eng._be.traj_h[a._s, a._i, : self._n()])

        @property
        def index(self) -> int:
            return int(self._a._snap.cursor[self._a._s, self._a._i])

        def reach_end(self) -> bool:
            return self.index >= self._n() - 1

    class _Obs:
        def __init__(self, agent: "Agent"):
            self._a = agent

        @property
        def ahead_vehicle(self):
            a = self._a
            if not a._snap.lead_valid[a._s, a._i]:
                return None
            return (float(a._snap.lead_dis[a._s, a._i]), float(a._snap.lead_v[a._s, a._i]))

    def __init__(self, eng: "ScenarioEngine", index: int):
        self._eng, self._s, self._i = eng, eng._s, index
        sb = eng._be.static
        self.id = int(sb.agent_id[self._s, index])
        self.length = float(sb.length[self._s, index])
        self.width = float(sb.width[self._s, index])
        self.dynamic_flag = bool(sb.dynamic[self._s, index])
        self.runtime = Agent.Runtime(self)
        self.history_states = Agent.HistoryState(self)
        self.ref_trajectory = Agent._Trajectory(self)
        self.obs = Agent._Obs(self)

    @property
    def _snap(self) -> _Snapshot:
        return self._eng._snap

    @property
    def heading(self) -> float:
        return self.runtime.heading

    @property
    def v(self) -> float:
        return self.runtime.v

    def is_offroad(self) -> bool:
        """reference: agent.py:482-494 (any agent, active or not)"""
        r = self._eng._rewards_for(self._i)
        return bool(r["terms"][4] != 0)

    def get_route(self) -> np.ndarray:
        """reference: agent.py:468-480 — next 10 waypoints in the ego frame, zero padded"""
        return self._eng._rewards_for(self._i)["route"].copy()

    def set_ref_trajectory(self, traj, vel, heading):
        """reference: agent.py:462-466"""
        be = self._eng._be
        be.set_ref_trajectory(np.array([self._s], np.int32), np.array([self._i], np.int32),
                              np.asarray(traj, np.float32)[None], np.asarray(vel, np.float32)[None],
                              np.asarray(heading, np.float32)[None])
        self._eng._root._invalidate()


class AgentManager:
    """reference: lcsim/entity/agent/manager.py"""

    def __init__(self, eng: "ScenarioEngine"):
        self._eng = eng
        sb = eng._be.static
        n = int(sb.n_agents[eng._s])
        self._list = [Agent(eng, i) for i in range(n)]
        self.agents: Dict[int, Agent] = {a.id: a for a in self._list}
        self.ads_id = self._list[int(sb.ads_index[eng._s])].id if n else None

    def get_agent_by_id(self, agent_id: int) -> Optional[Agent]:
        return self.agents.get(agent_id)

    def get_ads(self) -> Optional[Agent]:
        return self.agents.get(self.ads_id)

    @property
    def active_agents(self) -> List[Agent]:
        sn, s = self._eng._snap, self._eng._s
        return [self._list[i] for i in sn.active[s, : int(sn.n_active[s])]]

    @property
    def agents_waiting_list(self) -> List[Agent]:
        sn, s = self._eng._snap, self._eng._s
        order = self._eng._be.static.wait_order[s, : len(self._list)]
        return [self._list[i] for i in order if sn.status[s, i] == native.WAITING]

    def get_agent_bbox(self, ids: Sequence[int]) -> np.ndarray:
        sn, s = self._eng._snap, self._eng._s
        rows = []
        for _id in ids:
            a = self.agents.get(_id)
            if a is None:
                continue
            x, y, h, _ = sn.pose64[s, a._i]
            rows.append([x, y, a.length, a.width, h])
        return np.array(rows)

    def get_agent_ref_trajs(self, ids: Sequence[int], time_step: int) -> List[np.ndarray]:
        sb, s = self._eng._be.static, self._eng._s
        out = []
        for _id in ids:
            a = self.agents.get(_id)
            if a is None or sb.agent_type[s, a._i] != proto.AGENT_TYPE_VEHICLE:
                continue
            k = a.ref_trajectory.index
            out.append(a.ref_trajectory.xy[k: k + time_step])
        return out

    def check_collision(self, agent_id: int) -> bool:
        a = self.agents.get(agent_id)
        if a is None or a.runtime.status != Agent.Status.ACTIVE:
            return False
        return bool(self._eng._rewards_for(a._i)["terms"][3] != 0)

    def check_collision_all(self) -> np.ndarray:
        sn, s = self._eng._snap, self._eng._s
        n = int(sn.n_active[s])
        if n == 0:
            return np.array([])
        return sn.coll_count[s, sn.active[s, :n]].astype(np.int64)

    def get_collision_rate(self) -> float:
        c = self.check_collision_all()
        return 0.0 if len(c) == 0 else c.sum() / len(c)

    def get_offroad_rate(self) -> float:
        sn, s = self._eng._snap, self._eng._s
        n = int(sn.n_active[s])
        if n == 0:
            return 0.0
        return sn.offroad[s, sn.active[s, :n]].sum() / n


class ScenarioEngine:
    """The reference's ``Engine`` surface for scenario ``s`` of a batch."""

    def __init__(self, root: "Engine", s: int):
        self._root, self._s = root, s
        self.agent_manager = AgentManager(self)
        self._map = None  # (lane_manager, road_manager, junction_manager) views, built on first access
        self.motion_diffuser = None
        self.guidance = None
        self._pending: Dict[int, Action] = {}

    @property
    def _be(self) -> BatchEngine:
        return self._root._batch

    # ---- read-only views with the reference's manager surface (the engine itself works on packed arrays)
    def _map_views(self):
        if self._map is None:
            self._map = mapview.build(self._root._scenario(self._s))
        return self._map

    @property
    def lane_manager(self):
        return self._map_views()[0]

    @property
    def road_manager(self):
        return self._map_views()[1]

    @property
    def junction_manager(self):
        return self._map_views()[2]

    @property
    def _snap(self) -> _Snapshot:
        return self._root._snapshot

    def _rewards_for(self, agent_index: int) -> dict:
        return self._root._rewards(self._s, agent_index)

    # ---- clock (engine.py:38-43): float64 accumulation like the reference
    @property
    def current_step(self) -> int:
        return self._root._step_count[self._s]

    @property
    def current_time(self) -> float:
        return self._root._time[self._s]

    # ---- this scenario only (lcsim/envs/gym_warpper.py keeps one Engine per scenario and advances the current one)
    def reset(self):
        self._root._reset_scenarios([self._s])

    def step(self, actions: Dict[int, "Action"] = {}):
        self._root._step_scenarios({self._s: dict(actions)})

    def release_deep_models(self):
        self.motion_diffuser = None

    def load_deep_models(self):
        self.motion_diffuser = None

    def get_step_return(self, agent_id: int):
        """reference: engine.py:160-201"""
        am = self.agent_manager
        agent = am.get_agent_by_id(agent_id)
        assert agent is not None
        r = self._rewards_for(agent._i)
        if not agent.runtime.is_finished():
            enc = self._root.scene_encoder
            scene_emb = np.zeros(128) if enc is None else np.asarray(enc(self, agent))
            obs = {"scene_emb": scene_emb, "route": r["route"].copy()}
        else:
            obs = {"scene_emb": np.zeros(128), "route": np.zeros((10, 2))}
        cfg = self._root.reward_config
        fin = agent.runtime.is_finished()
        # like the reference, only the configured keys appear, plus the ones _compute_rewards always sets
        rewards = {k: 0.0 for k in cfg.keys()}
        for k, v in zip(native.REWARD_KEYS, r["terms"]):
            if k == "destination" and not fin:
                continue
            rewards[k] = float(v)
        reward = sum(rewards[k] * coef for k, coef in cfg.items())
        truncated = self.current_step >= self._root.total_steps or fin
        return obs, reward, bool(r["terminated"]), truncated, {"rewards": rewards}


class Engine(ScenarioEngine):
    """reference: lcsim/engine/engine.py:22-260"""

    def __new__(cls, config: dict):
        # a scenario of lane-position agents (LaneIDM, agent.py:121-131,167-171) goes to the city engine's face
        data_dir = config["input"]["data_dir"]
        if isinstance(data_dir, (str, os.PathLike)) and os.path.exists(os.path.join(data_dir, "agents.pb")):
            from . import city

            if city.is_city_dir(data_dir):
                from .city_facade import CityFacadeEngine

                return CityFacadeEngine(config)
        return super().__new__(cls)

    def __init__(self, config: dict):
        self.config = config
        step = config["control"]["step"]
        self.start_time = step["start"]
        self.total_steps = step["total"]
        self.step_interval = step["interval"]
        c = config["control"]
        self.enable_poly_emb = c["enable_poly_emb"]
        self.enable_plan = c["enable_plan"]
        if self.enable_plan:
            assert self.enable_poly_emb
        self.plan_interval = c["plan_interval"]
        self.plan_init_step = c["plan_init_step"]
        self.device = c["device"]
        self.allow_new_obj = c["allow_new_obj"]
        self.reactive = c["reactive"]
        self.no_static = c["no_static"]
        self.reward_config = config["reward"]
        self.scene_encoder = None
        data_dir = config["input"]["data_dir"]
        dirs = [data_dir] if isinstance(data_dir, (str, os.PathLike)) else list(data_dir)
        for d in dirs:  # same failure mode as open() in the reference (engine.py:90-99)
            for fn in ("map.pb", "agents.pb"):
                if not os.path.exists(os.path.join(d, fn)):
                    raise FileNotFoundError(os.path.join(d, fn))
        dev = 0
        if isinstance(self.device, str) and self.device.startswith("cuda:"):
            dev = int(self.device.split(":")[1])
        # map.pb / agents.pb -> packed arrays by the native loader (lcsim_b200_load_dirs, threaded; bit-identical to
        # packer.pack_dirs); the Python-side map objects behind lane_manager / junction_manager are parsed on first use
        backend = config.get("_backend") or native.TorchCudaBackend(dev)
        self._dirs = dirs
        self._scenario_arrays: Dict[int, packer.ScenarioArrays] = {}
        static, self._poly_centres, _ = native.load_dirs(backend.lib, dirs, start=float(self.start_time),
                                                         interval=float(self.step_interval), total_steps=int(self.total_steps))
        self._batch = BatchEngine(static, reactive=self.reactive, no_static=self.no_static,
                                  allow_new_obj=self.allow_new_obj, with_metrics=True, device=dev, backend=backend)
        self._snapshot = _Snapshot(self._batch)
        self._rew_cache: dict = {}
        self._step_count = [0] * static.S
        self._time = [self.start_time] * static.S
        ScenarioEngine.__init__(self, self, 0)
        self.scenarios: List[ScenarioEngine] = [self] + [ScenarioEngine(self, s) for s in range(1, static.S)]

    # ---- internals
    def _scenario(self, s: int) -> "packer.ScenarioArrays":
        if s not in self._scenario_arrays:
            self._scenario_arrays[s] = packer.scenarios_from_dirs([self._dirs[s]])[0]
        return self._scenario_arrays[s]

    def _invalidate(self):
        self._snapshot.invalidate()
        self._rew_cache = {}

    def _rewards(self, s: int, agent_index: int) -> dict:
        """Engine._compute_rewards / get_route / check_collision / is_offroad for (s, agent), batched over all
        scenarios in one launch and cached until the next tick."""
        key = ("ads",) if agent_index == int(self._batch.static.ads_index[s]) else (s, agent_index)
        if key not in self._rew_cache:
            be = self._batch
            agent = None
            if key != ("ads",):
                agent = be.static.ads_index.copy()
                agent[s] = agent_index
            r = be.rewards(agent)
            self._rew_cache[key] = {k: be.backend.to_numpy(v) for k, v in r.items()}
        c = self._rew_cache[key]
        return {k: v[s] for k, v in c.items()}

    def _reset_scenarios(self, which: Sequence[int]):
        """Engine.reset of the scenarios `which` only (masked reset launch)."""
        mask = np.zeros(self._batch.S, np.uint8)
        mask[list(which)] = 1
        self._batch.reset(mask)
        for s in which:
            self._step_count[s] = 0
            self._time[s] = self.start_time
            self.scenarios[s]._pending = {}
        self._invalidate()

    def _step_scenarios(self, actions_by_scenario: Dict[int, Dict[int, Action]]):
        """Engine.step(actions) of the given scenarios only (masked tick launch); the others keep state and clock."""
        be = self._batch
        mask = np.zeros(be.S, np.uint8)
        K = max((len(a) for a in actions_by_scenario.values()), default=0)
        aa = np.full((be.S, max(K, 1)), -1, np.int32)
        at = np.zeros((be.S, max(K, 1)), np.int32)
        ad = np.zeros((be.S, max(K, 1), 5))
        for s, acts in actions_by_scenario.items():
            mask[s] = 1
            am = self.scenarios[s].agent_manager
            for k, (aid, act) in enumerate(acts.items()):
                a = am.get_agent_by_id(aid)
                if a is not None:
                    aa[s, k], at[s, k], ad[s, k] = a._i, act.type.value, act._row()
        if K == 0:
            be.step(mask=mask)
        else:
            be.step(aa, at, ad, mask=mask)
        for s in actions_by_scenario:
            self._time[s] += self.step_interval
            self._step_count[s] += 1
        self._invalidate()

    def scenario_view(self, s: int) -> ScenarioEngine:
        """The reference's per-scenario Engine surface for scenario s whose reset()/step() touch that scenario only
        (``engine.scenarios[0]`` is this object, whose reset()/step() drive the whole batch)."""
        if not hasattr(self, "_views"):
            self._views = [ScenarioEngine(self, i) for i in range(self._batch.S)]
            for v in self._views:
                v.config, v.total_steps, v.step_interval = self.config, self.total_steps, self.step_interval
                v.reward_config, v.start_time = self.reward_config, self.start_time
        return self._views[s]

    # ---- reference API
    def reset(self):
        self._batch.reset()
        self._step_count = [0] * self._batch.S
        self._time = [self.start_time] * self._batch.S
        for sc in self.scenarios:
            sc._pending = {}
        self._invalidate()

    def update(self, actions: Dict[int, Action] = {}):
        """Update stage (engine.py:221-225).  The CUDA tick fuses update and prepare; the actions are held
        until prepare() launches it."""
        self._pending = dict(actions)

    def prepare(self):
        """Prepare stage (engine.py:203-219): launches the fused tick with the actions given to update()."""
        be = self._batch
        per_s = [sc._pending for sc in self.scenarios]
        K = max((len(p) for p in per_s), default=0)
        if K == 0:
            be.step()
        else:
            aa = np.full((be.S, K), -1, np.int32)
            at = np.zeros((be.S, K), np.int32)
            ad = np.zeros((be.S, K, 5))
            for s, pend in enumerate(per_s):
                am = self.scenarios[s].agent_manager
                for k, (aid, act) in enumerate(pend.items()):
                    a = am.get_agent_by_id(aid)
                    if a is None:
                        continue
                    aa[s, k], at[s, k], ad[s, k] = a._i, act.type.value, act._row()
            be.step(aa, at, ad)
        for sc in self.scenarios:
            sc._pending = {}
        self._invalidate()

    def step(self, actions: Union[Dict[int, Action], Sequence[Dict[int, Action]]] = {}):
        """reference: engine.py:145-158.  For a batch, ``actions`` may be a list with one dict per scenario."""
        if isinstance(actions, dict):
            self.update(actions)
        else:
            for sc, a in zip(self.scenarios, actions):
                sc._pending = dict(a)
        self.prepare()
        for s in range(self._batch.S):
            self._time[s] += self.step_interval
            self._step_count[s] += 1

    def render(self):
        raise NotImplementedError("rendering is outside the stepping path (SURVEY.md §2 row 8); feed the reference's "
                                  "plot utilities from agent_manager.get_agent_bbox()")

    def release_deep_models(self):
        self.motion_diffuser = None

    def load_deep_models(self):
        self.motion_diffuser = None

    def close(self):
        self._batch.close()
