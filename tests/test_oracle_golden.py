"""Pins the CPU oracle (oracle/lcsim_oracle.c) against golden vectors recorded from the real
reference engine (oracle/gen_golden.py).  Integers must be bit-exact at every tick; floats
are bit-identical in the container that generated the fixtures and are checked to 1e-9 here
(libm / BLAS builds may differ by an ulp between hosts)."""
import numpy as np
import pytest

from golden_util import Case, case_names
from oracle import oracle as orc

ATOL = 1e-9


def _check_tick(c, o, k, tol):
    ref = c.ref
    n = int(ref["n_active"][k])
    assert n == o.n_active[0], f"tick {k}: n_active"
    np.testing.assert_array_equal(ref["active"][k, :n], o.active[0, :n], err_msg=f"tick {k}: active list order")
    np.testing.assert_array_equal(ref["status"][k], o.status[0], err_msg=f"tick {k}: status")
    np.testing.assert_array_equal(ref["cursor"][k], o.cursor[0], err_msg=f"tick {k}: cursor")
    for nm in "xyhv":
        np.testing.assert_allclose(getattr(o, nm)[0], ref[nm][k], rtol=0, atol=tol, err_msg=f"tick {k}: {nm}")
    o.metrics()
    np.testing.assert_array_equal(ref["coll"][k], o.coll_count[0], err_msg=f"tick {k}: collision counts")
    np.testing.assert_array_equal(ref["offroad"][k], o.offroad[0], err_msg=f"tick {k}: offroad")
    np.testing.assert_array_equal(ref["nearest"][k], o.nearest_edge[0], err_msg=f"tick {k}: nearest edge idx")
    act = ref["active"][k, :n]
    np.testing.assert_array_equal(ref["lead_valid"][k][act], o.lead_valid[0][act], err_msg=f"tick {k}: lead")
    m = ref["lead_valid"][k].astype(bool)
    np.testing.assert_allclose(o.lead_dis[0][m], ref["lead_dis"][k][m], rtol=0, atol=tol)
    np.testing.assert_allclose(o.lead_v[0][m], ref["lead_v"][k][m], rtol=0, atol=tol)
    terms, term = o.rewards(0, c.ads)
    np.testing.assert_allclose(terms, ref["rewards"][k], rtol=0, atol=max(tol, 1e-9), equal_nan=True,
                               err_msg=f"tick {k}: reward terms")
    assert term == bool(ref["terminated"][k])
    assert o.check_collision(0, c.ads) == bool(ref["ads_collision"][k])
    assert o.is_offroad(0, c.ads)[0] == bool(ref["ads_offroad"][k])
    np.testing.assert_allclose(o.route(0, c.ads), ref["route"][k], rtol=0, atol=tol)


@pytest.mark.parametrize("name", case_names())
def test_oracle_matches_reference_golden(name):
    c = Case(name)
    o = orc.Oracle(c.static, reactive=c.reactive, no_static=c.no_static, allow_new_obj=c.allow_new_obj)
    # a float32 re-plan makes the reference evaluate a few expressions in float32 through numpy's own
    # float32 kernels; allow float32-ulp slack there
    tol = 2e-6 if c.replan is not None else ATOL
    _check_tick(c, o, 0, ATOL)
    for k in range(c.n_steps):
        if c.replan is not None and c.replan["tick"] == k:
            for a in range(c.static.A):
                if c.static.dynamic[0, a] and c.replan["mask"][a]:
                    o.set_ref_trajectory(0, a, c.replan["xy"][a], c.replan["vel"][a], c.replan["heading"][a])
        assert int(c.ref["n_updates"][k + 1]) == int(o.n_active[0])
        o.step(*c.actions_for(k))
        _check_tick(c, o, k + 1, tol)
    np.testing.assert_array_equal(c.ref["ring_valid"], o.ring_valid[0])
    np.testing.assert_allclose(o.ring[0], c.ref["ring"], rtol=0, atol=tol)
    assert int(c.ref["n_updates"].sum()) == int(o.agent_steps[0])


def test_survey_pins():
    """totals of the four bundled-scenario rollouts."""
    pins = {"womd_1a1b6e_replay": (3512, 3507, 328, 174), "womd_1a1b6e_reactive": (3740, 3757, 890, 143),
            "womd_1a0df8_replay": (2698, 2701, 300, 120), "womd_1a0df8_reactive": (2797, 2810, 238, 183)}
    for name, (upd, act, coll, off) in pins.items():
        r = Case(name).ref
        assert int(r["n_updates"].sum()) == upd
        assert int(r["n_active"][1:].sum()) == act
        assert int(r["coll"][1:].sum()) == coll
        assert int(r["offroad"][1:].sum()) == off


def test_np_sum_restatement():
    rng = np.random.default_rng(0)
    for n in list(range(0, 200)) + [257, 1000, 4097]:
        a = rng.uniform(0, 3, n)
        assert orc.np_sum(a) == float(np.sum(a)), n


def test_wrap_angle_restatement():
    rng = np.random.default_rng(1)
    for a in list(rng.uniform(-20, 20, 2000)) + [0.0, np.pi, -np.pi, 3 * np.pi, -3 * np.pi, 1e-300]:
        assert orc.wrap_angle(float(a)) == float((a + np.pi) % (2 * np.pi) - np.pi)
