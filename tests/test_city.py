"""City engine (BASELINE configs[3]) through the C ABI on both backends: per-tick parity with golden
vectors of the REAL reference engine on MOSS-shaped cities (tests/golden_city) and with the Python restatement
(oracle/city_oracle.py, itself pinned to those vectors) on a larger seeded city.  Integer outputs (active list and
order, status, lane ids, leader flags, collision counts, off-road flags) bit-exact, floats 1e-9."""
import numpy as np
import pytest

from city_util import CityCase, case_names, check_against_ref
from lcsim_b200 import city, native
from lcsim_b200.city_engine import CityEngine
from oracle.city_oracle import CityOracle


@pytest.fixture(scope="module", params=["hostemu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "cuda":
        return native.TorchCudaBackend(0)
    from hostemu import NumpyEmuBackend

    return NumpyEmuBackend()


def _getter(ce, st):
    tn = ce.backend.to_numpy

    def get(nm):
        if nm == "n_active":
            return ce.scalar("n_active")
        if nm == "lane_pbid":
            return st.lane_pbid[tn(ce.lane)]
        if nm in ("x", "y"):
            return tn(ce.xy)[:, 0 if nm == "x" else 1]
        if nm == "h":
            return tn(ce.heading)
        return tn(getattr(ce, nm))

    return get


@pytest.mark.parametrize("name", case_names())
def test_city_matches_reference_golden(name, backend):
    c = CityCase(name)
    ce = CityEngine(c.static, backend=backend)
    get = _getter(ce, c.static)
    ce.reset()
    check_against_ref(c.ref, 0, get)
    for k in range(c.n_ticks):
        ce.step()
        check_against_ref(c.ref, k + 1, get)
        np.testing.assert_array_equal(ce.backend.to_numpy(ce.is_lc), c.ref["is_lc"][k + 1], err_msg=f"tick {k + 1}: is_lc")
    assert ce.scalar("lc_required") == int(name.endswith("_lc")) and ce.scalar("lc_error") == 0 and ce.scalar("tick") == c.n_ticks
    np.testing.assert_array_equal(ce.backend.to_numpy(ce.is_lc), c.ref["is_lc"][-1])
    np.testing.assert_array_equal(ce.backend.to_numpy(ce.ring_count), c.ref["ring_valid"])
    np.testing.assert_allclose(ce.ring_ordered(), c.ref["ring"], rtol=0, atol=1e-9)
    st = ce.backend.to_numpy(ce.stats)
    assert int(st[1]) == int(c.ref["n_active"][1:].sum())
    assert int(st[2]) == int(c.ref["coll"][1:].sum()) and int(st[3]) == int(c.ref["offroad"][1:].sum())
    ce.close()


def _oracle_ref_view(o, st):
    return {"n_active": np.asarray([len(o.active)]), "active": np.asarray([o.active + [-1] * (o.N - len(o.active))]),
            "status": o.status[None], "lane": st.lane_pbid[o.lane][None], "s": o.s[None], "v": o.v[None],
            "x": o.xy[None, :, 0], "y": o.xy[None, :, 1], "h": o.h[None], "lead_valid": o.lead_valid[None],
            "lead_dis": o.lead_dis[None], "lead_v": o.lead_v[None], "coll": o.coll_count[None], "offroad": o.offroad[None]}


def test_city_matches_oracle_on_larger_city(backend):
    """6 x 6 junctions, 900 vehicles, dense departures: every tick against the restatement; several ticks in one call"""
    raw = city.make_city(seed=11, grid=6, n_agents=900, depart_window=8.0)
    st = city.pack_city(raw)
    ce = CityEngine(st, backend=backend)
    o = CityOracle(st)
    get = _getter(ce, st)
    check_against_ref(_oracle_ref_view(o, st), 0, get)
    for k in range(70):
        n = 1 if k % 5 else 3
        ce.step(n)
        for _ in range(n):
            o.step()
        check_against_ref(_oracle_ref_view(o, st), 0, get)
    assert ce.scalar("lc_required") == 0 and not o.lc_required and ce.scalar("lc_error") == 0
    assert int(ce.backend.to_numpy(ce.stats)[0]) == o.agent_steps
    ce.reset()  # Engine.reset: same state as a fresh engine
    o.reset()
    check_against_ref(_oracle_ref_view(o, st), 0, get)
    ce.close()


def test_city_lane_changes_match_oracle(backend):
    """movements that only some lanes offer + random end lanes: forced lane changes with their live reads of
    neighbours' speeds, 500 vehicles, against the restatement every tick"""
    # (seed chosen so that no vehicle is still changing lanes when it reaches a junction: the reference raises there)
    raw = city.make_city(seed=15, grid=5, n_agents=400, require_lc=True, max_hops=5, depart_window=6.0)
    st = city.pack_city(raw)
    ce = CityEngine(st, backend=backend)
    o = CityOracle(st)
    get = _getter(ce, st)
    n_lc = 0
    for k in range(140):
        ce.step()
        o.step()
        check_against_ref(_oracle_ref_view(o, st), 0, get)
        np.testing.assert_array_equal(ce.backend.to_numpy(ce.is_lc), o.is_lc, err_msg=f"tick {k}: is_lc")
        n_lc += int(o.is_lc.sum())
    assert n_lc > 500 and ce.scalar("lc_required") == 1 and ce.scalar("lc_error") == 0
    ce.close()


def test_city_through_facade(tmp_path, backend):
    """Engine(config) on a MOSS-shaped map.pb / agents.pb directory, driven like the reference (reset, step, manager and
    agent attributes), against the golden vectors of the real engine for the same city"""
    from lcsim_b200.engine import Engine

    c = CityCase("city_g2_a24")
    raw = city.make_city(seed=2, grid=2, n_agents=24, max_hops=5)  # the generator call that produced the golden case
    d = str(tmp_path / "city")
    city.write_city(d, raw)
    cfg = {"input": {"data_dir": d, "model_path": "none", "model_config_path": "none"},
           "control": {"step": {"start": 0, "total": 100000, "interval": 0.1}, "enable_poly_emb": False, "enable_plan": False,
                       "plan_interval": 80, "plan_init_step": 10, "device": "cuda:0", "allow_new_obj": True, "reactive": False,
                       "no_static": False, "guidance": False},
           "reward": {}, "render": {}, "_backend": backend}
    eng = Engine(cfg)
    eng.reset()
    am = eng.agent_manager
    ids = list(am.agents.keys())
    assert ids == [int(x) for x in c.static.agent_id] and am.ads_id == ids[int(c.static.ads_index)]
    ref = c.ref
    for k in range(120):
        eng.step()
        t = k + 1
        act = am.active_agents
        n = int(ref["n_active"][t])
        assert [ids.index(a.id) for a in act] == list(ref["active"][t, :n]) and eng.current_step == t
        np.testing.assert_array_equal(am.check_collision_all(), ref["coll"][t][ref["active"][t, :n]])
        for a in act[:: max(1, n // 5)]:
            i = ids.index(a.id)
            assert a.runtime.lane.id == int(ref["lane"][t, i]) and a.runtime.status.value == int(ref["status"][t, i])
            np.testing.assert_allclose([a.runtime.s, a.v, a.heading], [ref["s"][t, i], ref["v"][t, i], ref["h"][t, i]], rtol=0, atol=1e-9)
            np.testing.assert_allclose(a.runtime.xy, [ref["x"][t, i], ref["y"][t, i]], rtol=0, atol=1e-9)
            assert a.is_offroad() == bool(ref["offroad"][t, i])
            assert (a.obs.ahead_vehicle is not None) == bool(ref["lead_valid"][t, i])
        if n:
            assert am.get_collision_rate() == pytest.approx(ref["coll"][t].sum() / n)
            assert am.get_offroad_rate() == pytest.approx(ref["offroad"][t].sum() / n)
    eng.close()


def test_city_rejects_bad_static(backend):
    raw = city.make_city(seed=6, grid=2, n_agents=10)
    st = city.pack_city(raw)
    st.home_lane = st.home_lane.copy()
    st.home_lane[0] = st.n_lanes + 3
    with pytest.raises(native.NativeLibraryError, match="home_lane"):
        CityEngine(st, backend=backend)
