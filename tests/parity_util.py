"""Per-tick comparison of a BatchEngine (CUDA or host emulation) with the golden vectors / the oracle."""
import numpy as np

from lcsim_b200 import native


def _np(be, x):
    return be.backend.to_numpy(x)


def check_engine_vs_golden_tick(c, be, k, tol):
    """c: golden_util.Case, be: BatchEngine with S=1; compares the state after tick k."""
    ref = c.ref
    n = int(ref["n_active"][k])
    assert n == int(_np(be, be.n_active)[0]), f"tick {k}: n_active"
    np.testing.assert_array_equal(ref["active"][k, :n], _np(be, be.active)[0, :n], err_msg=f"tick {k}: active order")
    np.testing.assert_array_equal(ref["status"][k], _np(be, be.status)[0], err_msg=f"tick {k}: status")
    np.testing.assert_array_equal(ref["cursor"][k], _np(be, be.cursor)[0], err_msg=f"tick {k}: cursor")
    pose = _np(be, be.pose64)[0]
    for i, nm in enumerate("xyhv"):
        np.testing.assert_allclose(pose[:, i], ref[nm][k], rtol=0, atol=tol, err_msg=f"tick {k}: {nm}")
    np.testing.assert_array_equal(ref["coll"][k], _np(be, be.coll_count)[0], err_msg=f"tick {k}: collision counts")
    np.testing.assert_array_equal(ref["offroad"][k], _np(be, be.offroad)[0], err_msg=f"tick {k}: offroad")
    np.testing.assert_array_equal(ref["nearest"][k], _np(be, be.nearest_edge)[0], err_msg=f"tick {k}: nearest edge")
    if be.reactive:
        act = ref["active"][k, :n]
        np.testing.assert_array_equal(ref["lead_valid"][k][act], _np(be, be.lead_valid)[0][act], err_msg=f"tick {k}: lead")
        m = ref["lead_valid"][k].astype(bool)
        np.testing.assert_allclose(_np(be, be.lead_dis)[0][m], ref["lead_dis"][k][m], rtol=0, atol=tol)
        np.testing.assert_allclose(_np(be, be.lead_v)[0][m], ref["lead_v"][k][m], rtol=0, atol=tol)
    r = be.rewards()
    np.testing.assert_allclose(_np(be, r["terms"])[0], ref["rewards"][k], rtol=0, atol=max(tol, 1e-9), equal_nan=True,
                               err_msg=f"tick {k}: reward terms")
    assert bool(_np(be, r["terminated"])[0]) == bool(ref["terminated"][k]), f"tick {k}: terminated"
    assert (_np(be, r["terms"])[0][3] != 0) == bool(ref["ads_collision"][k]), f"tick {k}: ads collision"
    assert (_np(be, r["terms"])[0][4] != 0) == bool(ref["ads_offroad"][k]), f"tick {k}: ads offroad"
    np.testing.assert_allclose(_np(be, r["route"])[0], ref["route"][k], rtol=0, atol=tol, err_msg=f"tick {k}: route")


def run_case_against_golden(c, be, tol=1e-9, replan_tol=2e-6):
    tol_after = replan_tol if c.replan is not None else tol
    check_engine_vs_golden_tick(c, be, 0, tol)
    for k in range(c.n_steps):
        if c.replan is not None and c.replan["tick"] == k:
            sel = [a for a in range(c.static.A) if c.static.dynamic[0, a] and c.replan["mask"][a]]
            be.set_ref_trajectory(np.zeros(len(sel), np.int32), np.asarray(sel, np.int32), c.replan["xy"][sel],
                                  c.replan["vel"][sel], c.replan["heading"][sel])
        assert int(c.ref["n_updates"][k + 1]) == int(_np(be, be.n_active)[0])
        be.step(*c.actions_for(k))
        check_engine_vs_golden_tick(c, be, k + 1, tol_after)
    np.testing.assert_array_equal(c.ref["ring_valid"], _np(be, be.ring_count)[0])
    np.testing.assert_allclose(_np(be, be.ring_ordered())[0], c.ref["ring"], rtol=0, atol=tol_after)
    stats = _np(be, be.stats)[0]
    assert int(c.ref["n_updates"].sum()) == int(stats[0])
    assert int(c.ref["n_active"][1:].sum()) == int(stats[1])
    assert int(c.ref["coll"][1:].sum()) == int(stats[2])
    assert int(c.ref["offroad"][1:].sum()) == int(stats[3])


def compare_engine_with_oracle(be, o, tol, check_lead=True, near_duplicate_edges=None):
    """Whole-batch comparison of the current state of a BatchEngine with an oracle.Oracle (after o.metrics())."""
    A = be.A
    np.testing.assert_array_equal(_np(be, be.n_active), o.n_active, err_msg="n_active")
    na = o.n_active
    act_e, act_o = _np(be, be.active), o.active
    for s in range(be.S):
        np.testing.assert_array_equal(act_e[s, : na[s]], act_o[s, : na[s]], err_msg=f"scenario {s}: active order")
    np.testing.assert_array_equal(_np(be, be.status), o.status, err_msg="status")
    np.testing.assert_array_equal(_np(be, be.cursor), o.cursor, err_msg="cursor")
    pose = _np(be, be.pose64)
    for i, nm in enumerate("xyhv"):
        np.testing.assert_allclose(pose[..., i], getattr(o, nm), rtol=0, atol=tol, err_msg=nm)
    np.testing.assert_array_equal(_np(be, be.coll_count), o.coll_count, err_msg="collision counts")
    np.testing.assert_array_equal(_np(be, be.offroad), o.offroad, err_msg="offroad")
    ne, want = _np(be, be.nearest_edge), o.nearest_edge
    if near_duplicate_edges is None:
        np.testing.assert_array_equal(ne, want, err_msg="nearest edge")
    else:
        # Poses that went through sin/cos agree with the oracle to an ulp, not to the bit.  Where two
        # edge points coincide to 1e-9 m (joints of resampled polylines), that ulp may pick the other twin: accepted
        # here only if the two indices ARE such twins; everything else stays bit-exact.
        for s, a in np.argwhere(ne != want):
            e = near_duplicate_edges[s]
            assert ne[s, a] >= 0 and want[s, a] >= 0 and np.abs(e[ne[s, a]] - e[want[s, a]]).max() < 1e-9, \
                f"nearest edge of scenario {s} agent {a}: {ne[s, a]} != {want[s, a]}"
    if check_lead and be.reactive:
        activem = o.status == native.ACTIVE
        np.testing.assert_array_equal(_np(be, be.lead_valid)[activem], o.lead_valid[activem], err_msg="lead_valid")
        m = activem & (o.lead_valid > 0)
        np.testing.assert_allclose(_np(be, be.lead_dis)[m], o.lead_dis[m], rtol=0, atol=tol)
        np.testing.assert_allclose(_np(be, be.lead_v)[m], o.lead_v[m], rtol=0, atol=tol)
