"""lcsim_b200.envs: BicycleSingleEnv with the reference's constructor / spaces / reset-step protocol and worker
sharding (lcsim/envs/gym_warpper.py:12-92), and the batched BicycleVectorEnv, driven with the golden ADS actions of
`synth1_ads_bicycle_*` against what the REAL reference engine returned for the same actions (route, reward terms,
terminated), plus the read-only map-manager views."""
import numpy as np
import pytest

from golden_util import Case
from lcsim_b200 import native, synth
from lcsim_b200.envs import BicycleSingleEnv, BicycleVectorEnv

REWARD = {"forward": 2, "collision": 5, "smooth": 1, "destination": 1}  # the config the goldens were recorded with


@pytest.fixture(scope="module", params=["hostemu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "cuda":
        return native.TorchCudaBackend(0)
    from hostemu import NumpyEmuBackend

    return NumpyEmuBackend()


class EnvContext(dict):
    """stand-in for RLlib's EnvContext: a dict with worker_index / num_workers attributes"""
    worker_index, num_workers = 1, 1


def make_config(dataset, n, backend, reactive=False, cls=dict):
    return cls({
        "input": {"dataset_path": dataset, "num_scenarios": n, "model_path": "none", "model_config_path": "none"},
        "control": {"step": {"start": 0, "total": 91, "interval": 0.1}, "enable_poly_emb": False, "enable_plan": False,
                    "plan_interval": 80, "plan_init_step": 10, "device": "cuda:0", "allow_new_obj": True,
                    "release_idle": True, "reactive": reactive, "no_static": False, "guidance": False},
        "reward": dict(REWARD), "render": {}, "_backend": backend,
    })


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    root = tmp_path_factory.mktemp("dataset")
    # sorted listing: a_seed7 (other scenario), b_seed1 (the golden's scenario), c_seed9
    for name, seed, n in (("a_seed7", 7, 32), ("b_seed1", 1, 64), ("c_seed9", 9, 48)):
        synth.write_scenario(str(root / name), seed, n)
    return str(root)


def check_step(c, t, obs, reward, terminated, truncated, info):
    ref = c.ref
    terms = ref["rewards"][t]
    fin = int(ref["status"][t, c.ads]) in (native.IDLE, native.FINISHED)
    if not fin:
        np.testing.assert_allclose(obs["route"], ref["route"][t], rtol=0, atol=1e-9)
    else:
        assert not obs["route"].any()
    assert obs["scene_emb"].shape == (128,) and obs["route"].shape == (10, 2)
    assert bool(terminated) == bool(ref["terminated"][t])
    assert bool(truncated) == (t >= 91 or fin)
    want = sum(REWARD[k] * terms[i] for i, k in enumerate(native.REWARD_KEYS) if k in REWARD and (k != "destination" or fin))
    np.testing.assert_allclose(reward, want, rtol=0, atol=1e-8, equal_nan=True)
    for i, k in enumerate(native.REWARD_KEYS):
        if k in info["rewards"] and (k != "destination" or fin):
            np.testing.assert_allclose(info["rewards"][k], terms[i], rtol=0, atol=1e-9, equal_nan=True)


@pytest.mark.parametrize("reactive", [False, True])
def test_single_env_matches_reference(dataset, backend, reactive):
    c = Case("synth1_ads_bicycle_reactive" if reactive else "synth1_ads_bicycle_replay")
    env = BicycleSingleEnv(make_config(dataset, 3, backend, reactive))
    assert env.action_space.shape == (2,) and env.observation_space["route"].shape == (10, 2)
    assert len(env.engine_list) == 3 and env.cur_index == 0
    # gym_warpper.py:60-73: reset() resets the CURRENT scenario, returns its observation, then moves on to the next one
    obs, info = env.reset()
    assert env.cur_index == 1 and obs["route"].shape == (10, 2)
    # the env now sits on b_seed1, which nobody has reset since construction: Engine.__init__ state == reset state
    assert env.agent_id == int(c.ref["agent_ids"][c.ads])
    other = env.engine_list[0]
    other_state = other.agent_manager.get_ads().runtime.xy.copy()
    for k in range(c.n_steps):
        ret = env.step(c.ads_actions[k])
        check_step(c, k + 1, *ret)
        assert env.engine.current_step == k + 1
    # stepping scenario 1 left scenario 0 (and its clock) alone
    np.testing.assert_array_equal(other.agent_manager.get_ads().runtime.xy, other_state)
    assert other.current_step == 0 and env.engine_list[2].current_step == 0
    # a second pass over the same scenario after the round robin gives the same numbers (reset restores everything)
    env.reset()  # resets b_seed1, moves to c_seed9
    env.reset()  # resets c_seed9, moves to a_seed7
    env.reset()  # resets a_seed7, moves to b_seed1
    assert env.cur_index == 1 and env.engine.current_step == 0
    for k in range(10):
        check_step(c, k + 1, *env.step(c.ads_actions[k]))
    # map managers: the reference's read-only surface
    eng = env.engine
    j = eng.junction_manager.get_unique_junction()
    assert j.id == 300000000 and len(eng.road_manager.roads) == 0
    assert set(j.lane_ids) == set(eng.lane_manager.lanes.keys()) and len(j.lane_ids) > 100
    lane = eng.lane_manager.get_lane_by_id(j.lane_ids[0])
    assert lane.center_line.shape[1] == 2 and lane.length > 0 and lane.in_junction()
    env.close()


def test_worker_sharding(dataset, backend):
    """gym_warpper.py:36-47: scenario i goes to worker (i mod num_workers) + 1"""
    class W2(EnvContext):
        worker_index, num_workers = 2, 2

    env = BicycleSingleEnv(make_config(dataset, 3, backend, cls=W2))
    assert len(env.engine_list) == 1  # only b_seed1 (index 1)
    assert env.agent_id == int(Case("synth1_ads_bicycle_replay").ref["agent_ids"][-1])
    env.close()
    env = BicycleSingleEnv(make_config(dataset, 3, backend, cls=EnvContext))
    assert len(env.engine_list) == 3
    env.close()
    with pytest.raises(AssertionError):
        BicycleSingleEnv(make_config(dataset, 4, backend))  # num_scenarios > listing
    with pytest.raises(AssertionError):
        BicycleSingleEnv(make_config(dataset, 3, backend), render_mode="human")


def test_vector_env_matches_reference(dataset, backend):
    c = Case("synth1_ads_bicycle_replay")
    env = BicycleVectorEnv(make_config(dataset, 3, backend))
    assert env.num_envs == 3
    obs, _ = env.reset()
    assert obs["route"].shape == (3, 10, 2) and obs["scene_emb"].shape == (3, 128)
    rng = np.random.default_rng(0)
    for k in range(c.n_steps):
        acts = rng.uniform(-1, 1, (3, 2))
        acts[1] = c.ads_actions[k]
        obs, reward, terminated, truncated, info = env.step(acts)
        one_info = {"rewards": {key: v[1] for key, v in info["rewards"].items()}}
        check_step(c, k + 1, {"route": obs["route"][1], "scene_emb": obs["scene_emb"][1]}, reward[1], terminated[1],
                   truncated[1], one_info)
    # masked reset: env 1 restarts, the others keep their clocks
    env.reset(mask=np.array([0, 1, 0]))
    assert env.engine.scenarios[1].current_step == 0 and env.engine.scenarios[0].current_step == c.n_steps
    obs, reward, terminated, truncated, info = env.step(np.tile(c.ads_actions[0], (3, 1)))
    check_step(c, 1, {"route": obs["route"][1], "scene_emb": obs["scene_emb"][1]}, reward[1], terminated[1], truncated[1],
               {"rewards": {key: v[1] for key, v in info["rewards"].items()}})
    env.close()


def test_vector_env_observation_tensors(dataset, backend):
    """obs_neighbours registers the gather once (lcsim_b200_set_env_obs): after reset and after every step the scene
    encoder hook receives the tensors of the CURRENT state — the same bytes a fresh build_obs call produces."""
    kw = dict(k_agents=8, k_polys=16, radius=50.0, sort_by_distance=True, with_polys=False)
    env = BicycleVectorEnv(make_config(dataset, 3, backend), obs_neighbours=kw)
    tn = env._be.backend.to_numpy
    seen = []

    def encoder(engine, nb):
        seen.append({k: tn(v).copy() for k, v in nb.items()})
        return np.ones((3, 128))

    env.scene_encoder = encoder
    rng = np.random.default_rng(2)
    obs, _ = env.reset()
    for k in range(12):
        if k:
            obs, *_ = env.step(rng.uniform(-1, 1, (3, 2)))
        assert len(seen) == k + 1 and obs["scene_emb"].shape == (3, 128)
        fresh = env._be.build_obs(**kw)
        env._be.sync()
        for name, v in fresh.items():
            np.testing.assert_array_equal(seen[-1][name], tn(v), err_msg=f"{name} at step {k}")
    env.close()
