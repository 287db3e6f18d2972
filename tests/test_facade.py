"""The drop-in Python face (lcsim_b200/engine.py: Engine / AgentManager / Agent / Action with the reference's names)
driven the way the reference's callers drive it — notebook loop (examples/simulation.ipynb:105-112) and
BicycleSingleEnv.step (lcsim/envs/gym_warpper.py:78-87) — against the golden vectors of the real engine.
Scenario files are regenerated from the same seeds that produced the goldens (synth.write_scenario)."""
import numpy as np
import pytest

from golden_util import Case
from lcsim_b200 import native, synth
from lcsim_b200.engine import Action, Agent, Engine


@pytest.fixture(scope="module", params=["hostemu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "cuda":
        return native.TorchCudaBackend(0)
    from hostemu import NumpyEmuBackend

    return NumpyEmuBackend()


def make_config(data_dir, backend, **control):
    cfg = {
        "input": {"data_dir": data_dir, "model_path": "none", "model_config_path": "none"},
        "control": {"step": {"start": 0, "total": 91, "interval": 0.1}, "enable_poly_emb": False, "enable_plan": False,
                    "plan_interval": 80, "plan_init_step": 10, "device": "cuda:0", "allow_new_obj": True,
                    "reactive": False, "no_static": False, "guidance": False},
        "reward": {"forward": 2, "collision": 5, "smooth": 1, "destination": 1},  # lcsim/config/example.yml
        "render": {},
        "_backend": backend,
    }
    cfg["control"].update(control)
    return cfg


@pytest.mark.parametrize("name,seed,n,reactive", [("synth1_ads_bicycle_replay", 1, 64, False),
                                                  ("synth1_ads_bicycle_reactive", 1, 64, True)])
def test_env_loop_through_facade(tmp_path, backend, name, seed, n, reactive):
    c = Case(name)
    d = str(tmp_path / "scen")
    synth.write_scenario(d, seed, n)
    eng = Engine(make_config(d, backend, reactive=reactive))
    eng.reset()
    am = eng.agent_manager
    ids = [int(x) for x in c.ref["agent_ids"]]
    assert list(am.agents.keys()) == ids and am.ads_id == ids[c.ads]
    ref = c.ref
    for k in range(c.n_steps):
        data = Action.BicycleAction()
        data.init_with_normalized(float(c.ads_actions[k, 0]), float(c.ads_actions[k, 1]))
        eng.step({am.get_ads().id: Action(type=Action.ActionType.BICYCLE, data=data)})
        obs, reward, terminated, truncated, info = eng.get_step_return(am.ads_id)
        t = k + 1
        assert eng.current_step == t
        act = am.active_agents
        assert [ids.index(a.id) for a in act] == list(ref["active"][t, : ref["n_active"][t]])
        np.testing.assert_array_equal(am.check_collision_all(), ref["coll"][t][ref["active"][t, : len(act)]])
        ads = am.get_ads()
        np.testing.assert_allclose(ads.runtime.xy, [ref["x"][t, c.ads], ref["y"][t, c.ads]], rtol=0, atol=1e-9)
        assert ads.runtime.status == Agent.Status(int(ref["status"][t, c.ads]))
        assert ads.ref_trajectory.index == int(ref["cursor"][t, c.ads])
        assert bool(terminated) == bool(ref["terminated"][t])
        fin = ads.runtime.is_finished()
        assert bool(truncated) == (t >= 91 or fin)
        keys = ["forward", "route_progress", "smooth", "collision", "offroad", "destination"]
        for i, key in enumerate(keys):
            if key in info["rewards"]:
                np.testing.assert_allclose(info["rewards"][key], ref["rewards"][t][i], rtol=0, atol=1e-9, equal_nan=True)
        want = sum(eng.reward_config[key] * ref["rewards"][t][keys.index(key)] for key in eng.reward_config)
        np.testing.assert_allclose(reward, want, rtol=0, atol=1e-9)
        assert obs["scene_emb"].shape == (128,) and obs["route"].shape == (10, 2)
        if not fin:
            np.testing.assert_allclose(obs["route"], ref["route"][t], rtol=0, atol=1e-9)
            assert am.check_collision(am.ads_id) == bool(ref["ads_collision"][t])
            assert ads.is_offroad() == bool(ref["ads_offroad"][t])
        if k % 15 == 0:
            n_act = len(act)
            assert abs(am.get_collision_rate() - ref["coll"][t].sum() / max(n_act, 1)) < 1e-12
            assert abs(am.get_offroad_rate() - ref["offroad"][t].sum() / max(n_act, 1)) < 1e-12
            hs = ads.history_states
            assert hs.xyz.shape == (11, 3) and hs.valid.shape == (11, 1) and hs.type[0] == 1
    eng.close()


def test_facade_errors_like_the_reference(tmp_path, backend):
    with pytest.raises(FileNotFoundError):
        Engine(make_config(str(tmp_path / "missing"), backend))
    cfg = make_config(str(tmp_path), backend)
    del cfg["control"]["reactive"]
    with pytest.raises(KeyError):  # the reference indexes the dict directly (engine.py:59-86)
        Engine(cfg)
    with pytest.raises(AssertionError):
        Action(type=Action.ActionType.BICYCLE, data=Action.StateAction())


def test_batch_of_directories(tmp_path, backend):
    """data_dir as a list: every scenario is stepped by the same launch; engine.scenarios[i] is its facade."""
    dirs = []
    for i, (seed, n) in enumerate([(1, 64), (3, 96)]):
        d = str(tmp_path / f"s{i}")
        synth.write_scenario(d, seed, n)
        dirs.append(d)
    eng = Engine(make_config(dirs, backend))
    c = Case("synth3_nonew_replay")  # seed 3, 96 agents, but recorded with allow_new_obj=False -> compare tick 0 only
    assert len(eng.scenarios) == 2
    act = eng.scenarios[1].agent_manager.active_agents
    assert len(act) == int(c.ref["n_active"][0])
    for _ in range(5):
        eng.step([{}, {}])
    assert eng.scenarios[1].current_step == 5 and abs(eng.scenarios[0].current_time - 0.5) < 1e-12
    eng.close()
