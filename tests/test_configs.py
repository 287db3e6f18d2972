"""BASELINE.json configs[2] and configs[4] as parity cases (configs[0]/[1]: test_parity.py, configs[3]: test_city.py).

configs[2] "RL env rollout: parallel waymo_ppo envs with observation + reward build": every scenario is one
BicycleSingleEnv (lcsim/envs/gym_warpper.py:60-92, config experiments/rl/configs/waymo_ppo.yml:10-41: reactive False,
total 90; the reactive variant is run too); each tick a
policy action per env goes in through lcsim_b200_env_step, reward terms / flags / route come back, the neighbour
gather feeds the observation.  configs[4] "diffusion-controlled rollout": every plan_interval ticks the dynamic agents
of every scenario receive new float32 (80, .) reference trajectories (Engine._plan_trajectory, engine.py:286-297 —
random tensors stand in for the denoiser) and the batch keeps stepping."""
import numpy as np
import pytest

from lcsim_b200 import native, workload
from lcsim_b200.batch import BatchEngine
from oracle import oracle as orc
from parity_util import compare_engine_with_oracle

WAYMO_PPO_REWARD = {"forward": 0.1, "collision": 10, "offroad": 5, "smooth": 2, "destination": 1}  # waymo_ppo.yml:36-41


@pytest.fixture(scope="module", params=["hostemu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "cuda":
        return native.TorchCudaBackend(0)
    from hostemu import NumpyEmuBackend

    return NumpyEmuBackend()


@pytest.mark.parametrize("reactive", [False, True])
def test_rl_env_rollout_matches_oracle(backend, reactive):
    S = 6
    sb = workload.synthetic_batch(400, S, 128, workers=1, total_steps=90)
    be = BatchEngine(sb, reactive=reactive, with_metrics=True, backend=backend)
    o = orc.Oracle(sb, reactive=reactive)
    rng = np.random.default_rng(3)
    out = np.zeros(S, native.ENV_OUT_DTYPE)
    ads = sb.ads_index
    done = np.zeros(S, bool)
    for k in range(90):
        act = np.stack([rng.uniform(-1, 1, S), rng.uniform(-1, 1, S)], axis=1)
        be.env_step(act, out, coef=WAYMO_PPO_REWARD, sync=True)
        aa = ads.reshape(S, 1).astype(np.int32)
        at = np.full((S, 1), native.ACT_BICYCLE, np.int32)
        ad = np.zeros((S, 1, 5))
        ad[:, 0, 0], ad[:, 0, 1] = act[:, 0] * 6.0, act[:, 1] * 0.3  # BicycleAction.init_with_normalized (type.py:100-102)
        o.step(aa, at, ad)
        o.metrics()
        compare_engine_with_oracle(be, o, 1e-9, near_duplicate_edges=sb.edge_xy)
        for s in range(S):
            terms, terminated = o.rewards(s, int(ads[s]))
            np.testing.assert_allclose(out["terms"][s], terms, rtol=0, atol=1e-9, equal_nan=True, err_msg=f"tick {k} env {s}")
            assert bool(out["terminated"][s]) == terminated
            fin = o.status[s, ads[s]] in (orc.IDLE, orc.FINISHED)
            assert bool(out["finished"][s]) == fin and bool(out["truncated"][s]) == (k + 1 >= sb.total_steps or fin)
            want = sum(WAYMO_PPO_REWARD[key] * terms[i] for i, key in enumerate(native.REWARD_KEYS) if key in WAYMO_PPO_REWARD)
            np.testing.assert_allclose(out["reward"][s], want, rtol=0, atol=1e-8, equal_nan=True)
            if not fin:
                np.testing.assert_allclose(out["route"][s], o.route(s, int(ads[s])), rtol=0, atol=1e-9)
            done[s] |= fin
        if k % 30 == 29:  # envs whose episode ended are reset, the others keep running (gym reset of single envs)
            mask = done.astype(np.uint8)
            be.reset(mask)
            if mask.any():
                o.reset(np.nonzero(mask)[0])
                o.metrics()
            done[:] = False
            compare_engine_with_oracle(be, o, 1e-9, near_duplicate_edges=sb.edge_xy)
    be.close()


@pytest.mark.parametrize("reactive", [False, True])
def test_diffusion_replan_rollout_matches_oracle(backend, reactive):
    S, T_PLAN, PLAN_INIT, PLAN_INTERVAL = 4, 80, 10, 20  # example.yml: plan_init_step 10; shorter interval to re-plan 4x
    sb = workload.synthetic_batch(500, S, 128, workers=1)
    be = BatchEngine(sb, reactive=reactive, with_metrics=True, backend=backend)
    o = orc.Oracle(sb, reactive=reactive)
    rng = np.random.default_rng(9)
    tn = be.backend.to_numpy
    for k in range(91):
        if k >= PLAN_INIT and (k - PLAN_INIT) % PLAN_INTERVAL == 0:
            sc, ag, xy, vel, hd = [], [], [], [], []
            for s in range(S):
                for a in range(sb.n_agents[s]):
                    if not sb.dynamic[s, a] or o.status[s, a] != orc.ACTIVE:
                        continue  # engine.py:293-297: dynamic agents of the junction's active set
                    speed = rng.uniform(0.0, 12.0)
                    h0 = o.h[s, a] + rng.normal(0, 0.05)
                    hh = (h0 + np.cumsum(rng.normal(0, 0.01, T_PLAN))).astype(np.float32)
                    step = np.stack([np.cos(hh), np.sin(hh)], axis=1) * speed * 0.1
                    p = (np.asarray([o.x[s, a], o.y[s, a]]) + np.cumsum(step, axis=0)).astype(np.float32)
                    v = (np.stack([np.cos(hh), np.sin(hh)], axis=1) * speed).astype(np.float32)
                    sc.append(s); ag.append(a); xy.append(p); vel.append(v); hd.append(hh)
                    o.set_ref_trajectory(s, a, p, v, hh)
            assert len(sc) > 20
            be.set_ref_trajectory(np.asarray(sc, np.int32), np.asarray(ag, np.int32), np.stack(xy), np.stack(vel), np.stack(hd))
        be.step()
        o.step()
        o.metrics()
        compare_engine_with_oracle(be, o, 2e-6 if k >= PLAN_INIT else 1e-9, near_duplicate_edges=sb.edge_xy)
    np.testing.assert_array_equal(tn(be.stats)[:, 0], o.agent_steps)
    be.close()
