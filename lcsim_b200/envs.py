"""``lcsim.envs`` on the CUDA engine: ``BicycleSingleEnv`` with the reference's constructor, spaces, scenario
sharding and reset / step protocol (lcsim/envs/gym_warpper.py:12-92), and ``BicycleVectorEnv``, the batched form that
steps every scenario of the worker with one ``lcsim_b200_env_step`` call per tick.

The reference builds one ``Engine`` per scenario of the worker and advances only the current one.  Here the worker's
scenarios live in ONE batch on the GPU (one upload, one handle); ``engine_list[i]`` are per-scenario views of it whose
``reset()`` / ``step()`` launch masked ticks, so the single-env protocol (reset -> next scenario, round robin) is
unchanged for RLlib-style callers.

gymnasium is optional: without it ``spaces.Box`` / ``spaces.Dict`` are minimal stand-ins with the attributes callers
read (``low, high, shape, dtype, sample(), contains()``).
"""
from __future__ import annotations

import copy
import os
from typing import List, Optional

import numpy as np

from . import native
from .engine import Action, Engine, ScenarioEngine

try:  # pragma: no cover - depends on the environment
    import gymnasium as gym
    from gymnasium import spaces
except ImportError:  # minimal stand-ins
    class _Space:
        def __init__(self, shape=None, dtype=None):
            self.shape, self.dtype = shape, dtype

    class _Box(_Space):
        def __init__(self, low, high, shape, dtype=float):
            super().__init__(tuple(shape), np.dtype(dtype))
            self.low = np.full(shape, low, self.dtype)
            self.high = np.full(shape, high, self.dtype)
            self._rng = np.random.default_rng()

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    class _Dict(_Space):
        def __init__(self, spaces_):
            super().__init__()
            self.spaces = dict(spaces_)

        def __getitem__(self, k):
            return self.spaces[k]

        def sample(self):
            return {k: v.sample() for k, v in self.spaces.items()}

        def contains(self, x) -> bool:
            return isinstance(x, dict) and set(x) == set(self.spaces) and all(self.spaces[k].contains(v) for k, v in x.items())

    class _Spaces:
        Box, Dict = _Box, _Dict

    class _Gym:
        class Env:
            metadata: dict = {}

    gym, spaces = _Gym, _Spaces


def _worker_scenarios(config) -> List[str]:
    """the worker's share of the dataset, exactly as gym_warpper.py:32-47: sorted listing, first num_scenarios,
    round robin over the workers (worker_index is 1-based; 1 of 1 without RLlib's EnvContext)"""
    scenario_path_list = sorted(os.listdir(config["input"]["dataset_path"]))
    num_scenarios = config["input"]["num_scenarios"]
    assert num_scenarios <= len(scenario_path_list)
    try:
        worker_index, num_workers = config.worker_index, config.num_workers
    except AttributeError:
        worker_index, num_workers = 1, 1
    scenario_path_list = scenario_path_list[:num_scenarios]
    return [scenario_path_list[i] for i in range(worker_index - 1, len(scenario_path_list), num_workers)]


def _spaces():
    action_space = spaces.Box(low=-1.0, high=1.0, shape=(2,), dtype=float)
    observation_space = spaces.Dict({
        "scene_emb": spaces.Box(low=-100.0, high=100.0, shape=(128,), dtype=float),
        "route": spaces.Box(low=-100.0, high=100.0, shape=(10, 2), dtype=float),
    })
    return action_space, observation_space


def _batch_engine_config(config, data_dirs):
    backend = config.get("_backend")  # a live backend object (tests) is shared, not copied
    cfg = copy.deepcopy({k: v for k, v in dict(config).items() if k != "_backend"})
    if backend is not None:
        cfg["_backend"] = backend
    cfg["input"] = dict(cfg["input"])
    cfg["input"]["data_dir"] = [os.path.join(config["input"]["dataset_path"], d) for d in data_dirs]
    return cfg


class BicycleSingleEnv(gym.Env):
    """reference: lcsim/envs/gym_warpper.py:12-92"""

    metadata = {"render_modes": ["rgb_array"]}
    agent_ids: List[int]
    engine_list: List[ScenarioEngine]
    cur_index: int

    def __init__(self, config: dict, render_mode=None):
        assert render_mode is None or render_mode in self.metadata["render_modes"]
        self.render_mode = render_mode
        self.action_space, self.observation_space = _spaces()
        self.release_idle = config["control"]["release_idle"]
        data_dirs = _worker_scenarios(config)
        # one batch on the GPU for all scenarios of this worker; engine_list holds its per-scenario views
        self._batch_root = Engine(_batch_engine_config(config, data_dirs))
        self.engine_list = [self._batch_root.scenario_view(i) for i in range(len(data_dirs))]
        self.agent_ids = [e.agent_manager.ads_id for e in self.engine_list]
        if self.release_idle:
            for e in self.engine_list:
                e.release_deep_models()
        self.cur_index = 0
        self.engine = self.engine_list[self.cur_index]
        if self.release_idle:
            self.engine.load_deep_models()
        self.agent_id = self.agent_ids[self.cur_index]

    def reset(self, seed=None, options=None):
        self.engine.reset()
        obs, _, _, _, info = self.engine.get_step_return(self.agent_id)
        if self.release_idle:
            self.engine.release_deep_models()
        # move to the next scenario
        self.cur_index = (self.cur_index + 1) % len(self.engine_list)
        self.engine = self.engine_list[self.cur_index]
        if self.release_idle:
            self.engine.load_deep_models()
        self.agent_id = self.agent_ids[self.cur_index]
        return obs, info

    def step(self, action):
        data = Action.BicycleAction()
        data.init_with_normalized(action[0], action[1])
        a = Action(type=Action.ActionType.BICYCLE, data=data)
        action_dict = {self.engine.agent_manager.get_ads().id: a}
        self.engine.step(action_dict)
        return self.engine.get_step_return(self.agent_id)

    def render(self):
        if self.render_mode is None:
            return
        return self.engine.render()

    def close(self):
        self._batch_root.close()


class BicycleVectorEnv:
    """All scenarios of the worker as parallel envs: one ``lcsim_b200_env_step`` per tick for the whole batch
    (BASELINE configs[2]).  ``step(actions)`` takes (S,2) normalised (acc, steer) and returns
    ``(obs, reward (S,), terminated (S,), truncated (S,), info)`` with ``obs = {"scene_emb": (S,128), "route": (S,10,2)}``
    like S calls of BicycleSingleEnv.step; envs whose episode ended are reset by the next ``reset(mask)`` (or at once
    with ``auto_reset=True``, gymnasium's vector convention: the returned observation is then the first of the new
    episode and ``info["final_observation"]`` holds the last of the old one)."""

    metadata = {"render_modes": []}

    def __init__(self, config: dict, auto_reset: bool = False, obs_neighbours: Optional[dict] = None):
        self.single_action_space, self.single_observation_space = _spaces()
        data_dirs = _worker_scenarios(config)
        self.engine = Engine(_batch_engine_config(config, data_dirs))
        self.num_envs = len(data_dirs)
        self.auto_reset = bool(auto_reset)
        self.reward_config = dict(config["reward"])
        self._be = self.engine._batch
        self._out = np.zeros(self.num_envs, native.ENV_OUT_DTYPE)
        self.scene_encoder = None  # callable(engine, neighbour tensors) -> (S,128); zeros without a model
        self.obs_neighbours = obs_neighbours  # kwargs of BatchEngine.build_obs, or None: no neighbour gather
        # the gather is registered once: every env_step then leaves the observation of the new state in these tensors
        self._nb = self._be.set_env_obs(**obs_neighbours) if obs_neighbours is not None else None
        self._nb_fresh = False

    def _obs(self):
        fin = self._out["finished"].astype(bool)
        route = self._out["route"].copy()
        route[fin] = 0.0  # engine.py:189-192
        emb = np.zeros((self.num_envs, 128))
        if self._nb is not None:
            nb = self._nb if self._nb_fresh else self._be.build_env_obs()  # (after a reset no env_step built it)
            self._nb_fresh = False
            if self.scene_encoder is not None:
                emb = np.asarray(self.scene_encoder(self.engine, nb))
                emb[fin] = 0.0
        return {"scene_emb": emb, "route": route}

    def _read(self):
        """reward / flags of the current state without stepping (Engine.get_step_return after reset)"""
        r = self._be.rewards(coef=self.reward_config)
        tn = self._be.backend.to_numpy
        self._out["reward"], self._out["terms"] = tn(r["reward"]), tn(r["terms"])
        self._out["route"] = tn(r["route"])
        self._out["terminated"], self._out["truncated"] = tn(r["terminated"]), tn(r["truncated"])
        st = tn(self._be.status)[np.arange(self.num_envs), self._be.static.ads_index]
        self._out["finished"] = (st == native.IDLE) | (st == native.FINISHED)

    def reset(self, mask=None, seed=None, options=None):
        if mask is None:
            self.engine.reset()
        else:
            self.engine._reset_scenarios([int(i) for i in np.nonzero(np.asarray(mask))[0]])
        self._read()
        return self._obs(), {}

    def step(self, actions):
        actions = np.ascontiguousarray(actions, np.float64).reshape(self.num_envs, 2)
        self._be.env_step(actions, self._out, coef=self.reward_config, sync=True)
        self._nb_fresh = True
        for s in range(self.num_envs):
            self.engine._time[s] += self.engine.step_interval
            self.engine._step_count[s] += 1
        self.engine._invalidate()
        obs = self._obs()
        reward = self._out["reward"].copy()
        terminated, truncated = self._out["terminated"].astype(bool), self._out["truncated"].astype(bool)
        info = {"rewards": {k: self._out["terms"][:, i].copy() for i, k in enumerate(native.REWARD_KEYS)}}
        done = terminated | truncated
        if self.auto_reset and done.any():
            info["final_observation"] = obs
            obs, _ = self.reset(mask=done)
        return obs, reward, terminated, truncated, info

    def close(self):
        self.engine.close()
