"""Parity of the stepping path with the reference, through the C ABI.

Every test runs on two backends:
  * ``cuda``    (marked gpu)  — the sm_100a library on a B200: the parity tests proper;
  * ``hostemu`` (CPU tier)    — the same kernel text compiled as plain C++ (tests/hostemu), so the control
    flow, the ABI and the host-side preparation are exercised without a GPU.
Checkers: the golden vectors recorded from the real reference engine (tests/golden, oracle/gen_golden.py)
and the C oracle (oracle/lcsim_oracle.c, itself pinned to those vectors by test_oracle_golden.py).

Bar (BASELINE.json north_star): integer outputs bit-exact; float64 master poses within 1e-4 m / 1e-5 rad
after the 91-step rollout.  The asserted tolerances below are far tighter (1e-9 replay/IDM, 2e-6 after a
float32 re-plan, where the reference itself computes in float32).
"""
import numpy as np
import pytest

from golden_util import Case, case_names
from lcsim_b200 import native, workload
from lcsim_b200.batch import BatchEngine
from oracle import oracle as orc
from parity_util import compare_engine_with_oracle, run_case_against_golden

POSE_TOL = 1e-9  # contract: 1e-4 m / 1e-5 rad


@pytest.fixture(scope="module", params=["hostemu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "cuda":
        return native.TorchCudaBackend(0)
    from hostemu import NumpyEmuBackend

    return NumpyEmuBackend()


@pytest.fixture(scope="module")
def synth8():
    return workload.synthetic_batch(100, 8, 128, workers=1)


@pytest.mark.parametrize("name", case_names())
def test_matches_reference_golden(name, backend):
    c = Case(name)
    be = BatchEngine(c.static, reactive=c.reactive, no_static=c.no_static, allow_new_obj=c.allow_new_obj,
                     with_metrics=True, backend=backend)
    run_case_against_golden(c, be, tol=POSE_TOL)
    be.close()


@pytest.mark.parametrize("reactive", [False, True])
def test_synthetic_batch_matches_oracle_every_tick(backend, synth8, reactive):
    """8 WOMD-shaped scenarios x 128 agents (BASELINE configs[1] shape), 91 ticks, every tick compared."""
    sb = synth8
    be = BatchEngine(sb, reactive=reactive, with_metrics=True, backend=backend)
    o = orc.Oracle(sb, reactive=reactive)
    o.metrics()
    compare_engine_with_oracle(be, o, POSE_TOL)
    for k in range(91):
        be.step()
        o.step()
        o.metrics()
        compare_engine_with_oracle(be, o, POSE_TOL)
    st = be.backend.to_numpy(be.stats)
    np.testing.assert_array_equal(st[:, 0], o.agent_steps)
    tot = be.backend.to_numpy(be.episode_stats())
    assert int(tot[0]) == int(o.agent_steps.sum())
    np.testing.assert_allclose(be.backend.to_numpy(be.ring_ordered()), o.ring, rtol=0, atol=POSE_TOL)
    np.testing.assert_array_equal(be.backend.to_numpy(be.ring_count), o.ring_valid)
    be.close()


@pytest.mark.parametrize("depth", ["1", "2", "4"])
@pytest.mark.parametrize("reactive", [False, True])
def test_rollout_launch_matches_oracle(backend, synth8, reactive, depth, monkeypatch):
    """lcsim_b200_rollout_log: all ticks of all scenarios in ONE launch (tick items and metric items of a run-time ticket
    queue, metric items lagging behind over a snapshot ring of the given depth).  The end state equals the oracle's
    after the same number of steps, the per-tick log equals the oracle's per-tick sums, and a rollout split into
    uneven chunks (1 + 17 + 73 ticks) lands on bit-identical state."""
    monkeypatch.setenv("LCSIM_B200_SNAP_DEPTH", depth)  # read by lcsim_b200_create
    sb = synth8
    o = orc.Oracle(sb, reactive=reactive)
    want = np.zeros((91, sb.S, 4), np.int64)
    for k in range(91):
        want[k, :, 0] = o.n_active
        o.step()
        o.metrics()
        want[k, :, 1] = o.n_active
        want[k, :, 2] = o.coll_count.sum(axis=1)
        want[k, :, 3] = o.offroad.sum(axis=1)
    be = BatchEngine(sb, reactive=reactive, with_metrics=True, backend=backend)
    tn = be.backend.to_numpy
    log = tn(be.rollout(91, log=True))
    np.testing.assert_array_equal(log, want)
    compare_engine_with_oracle(be, o, POSE_TOL)
    np.testing.assert_allclose(tn(be.ring_ordered()), o.ring, rtol=0, atol=POSE_TOL)
    np.testing.assert_array_equal(tn(be.stats)[:, 0], o.agent_steps)
    keys = ("pose64", "pose32", "status", "cursor", "active", "coll_count", "nearest_edge", "offroad", "ring", "ring_head",
            "stats", "lead_valid", "lead_dis", "lead_v")
    snap = {k: tn(getattr(be, k)).copy() for k in keys}
    be.reset()
    parts = [tn(be.rollout(n, log=True)) for n in (1, 17, 73)]
    np.testing.assert_array_equal(np.concatenate(parts), want)
    for k, v in snap.items():
        np.testing.assert_array_equal(tn(getattr(be, k)), v, err_msg=k)
    be.close()


def test_rollout_host_buffers(backend, synth8):
    """lcsim_b200_rollout_host: reset + 91 ticks + read-back into host memory equals the device views and the log"""
    sb = synth8.scenario_slice(0, 4)
    be = BatchEngine(sb, reactive=False, with_metrics=True, backend=backend)
    tn = be.backend.to_numpy
    want_log = tn(be.rollout(91, log=True)).copy()
    Ap = be.A_stride
    out = dict(tick_log=np.zeros((91, sb.S, 4), np.int32), pose32=np.zeros((sb.S, Ap, 4), np.float32),
               status=np.zeros((sb.S, Ap), np.int32), coll_count=np.zeros((sb.S, Ap), np.int32),
               nearest_edge=np.zeros((sb.S, Ap), np.int32), offroad=np.zeros((sb.S, Ap), np.uint8), stats=np.zeros(8, np.int64))
    be.rollout(7)  # dirty state: rollout_host resets first
    be.rollout_host(91, out)
    np.testing.assert_array_equal(out["tick_log"], want_log)
    for k in ("pose32", "status", "coll_count", "nearest_edge", "offroad"):
        np.testing.assert_array_equal(out[k][:, : be.A], tn(getattr(be, k)), err_msg=k)
    np.testing.assert_array_equal(out["stats"], tn(be.stats).sum(axis=0))
    assert out["stats"][0] == want_log[:, :, 0].sum()
    # masked: only scenario 2 restarts, the others continue from tick 91
    mask = np.array([0, 0, 1, 0], np.uint8)
    be.rollout_host(5, dict(stats=out["stats"]), reset_mask=mask)
    assert tn(be.tick).tolist() == [96, 96, 5, 96]
    be.close()


def test_oversized_boxes_collision_counts(backend, synth8):
    """boxes larger than a cell of the collision grid (articulated trucks, 18-30 m) take the all-partners path;
    far-away agents (external STATE actions outside the grid) are clamped into border cells"""
    import dataclasses

    sb = synth8.scenario_slice(0, 3)
    length, width = sb.length.copy(), sb.width.copy()
    rng = np.random.default_rng(21)
    for s in range(sb.S):
        big = rng.choice(sb.A, 12, replace=False)
        length[s, big] = rng.uniform(18.0, 30.0, 12)
        width[s, big] = rng.uniform(2.5, 9.0, 12)
    sb = dataclasses.replace(sb, length=length, width=width)
    be = BatchEngine(sb, reactive=False, with_metrics=True, backend=backend)
    o = orc.Oracle(sb, reactive=False)
    total = 0
    for k in range(60):
        aa = np.full((sb.S, 2), -1, np.int32)
        at = np.zeros((sb.S, 2), np.int32)
        ad = np.zeros((sb.S, 2, 5))
        if k % 7 == 3:  # two agents jump far outside the 256 m grid, next to each other
            aa[:, 0], aa[:, 1] = 5, 9
            at[:] = native.ACT_STATE
            ad[:, 0, :2] = [900.0, -700.0]
            ad[:, 1, :2] = [901.0, -700.5]
        be.step(aa, at, ad)
        o.step(aa, at, ad)
        o.metrics()
        compare_engine_with_oracle(be, o, POSE_TOL)
        total += int(o.coll_count.sum())
    assert total > 500
    be.close()


def test_external_actions_all_types(backend, synth8):
    """BICYCLE / DELTA / STATE actions for several agents per scenario (policy-action mode, agent.py:377-416)."""
    sb = synth8.scenario_slice(0, 4)
    rng = np.random.default_rng(7)
    for reactive in (False, True):
        be = BatchEngine(sb, reactive=reactive, with_metrics=True, backend=backend)
        o = orc.Oracle(sb, reactive=reactive)
        K = 5
        for k in range(40):
            aa = rng.integers(-1, sb.A, (sb.S, K)).astype(np.int32)
            aa[:, 0] = sb.ads_index
            at = rng.integers(0, 4, (sb.S, K)).astype(np.int32)
            ad = np.zeros((sb.S, K, 5))
            ad[..., 0] = rng.uniform(-6, 6, (sb.S, K))
            ad[..., 1] = rng.uniform(-0.3, 0.3, (sb.S, K))
            ad[..., 2] = rng.uniform(-0.2, 0.2, (sb.S, K))
            st = at == native.ACT_STATE
            ad[st, 0:2] = rng.uniform(-90, 90, (int(st.sum()), 2))
            ad[st, 2:4] = rng.uniform(-10, 10, (int(st.sum()), 2))
            ad[st, 4] = rng.uniform(-7, 7, int(st.sum()))
            dl = at == native.ACT_DELTA
            ad[dl, 0:2] = rng.uniform(-1, 1, (int(dl.sum()), 2))
            # the oracle applies the LAST matching entry; make agents unique per scenario to avoid ambiguity
            for s in range(sb.S):
                _, first = np.unique(aa[s], return_index=True)
                dup = np.ones(K, bool)
                dup[first] = False
                aa[s, dup] = -1
            be.step(aa, at, ad)
            o.step(aa, at, ad)
            o.metrics()
            compare_engine_with_oracle(be, o, POSE_TOL)
        be.close()


def test_reset_is_idempotent_and_masked(backend, synth8):
    sb = synth8.scenario_slice(0, 4)
    be = BatchEngine(sb, reactive=True, with_metrics=True, backend=backend)
    tn = be.backend.to_numpy
    be.rollout(30)
    snap = {k: tn(getattr(be, k)).copy() for k in ("pose64", "status", "cursor", "active", "coll_count", "ring")}
    mask = np.array([1, 0, 1, 0], np.uint8)
    be.reset(mask)
    assert tn(be.tick).tolist() == [0, 30, 0, 30]
    for k, v in snap.items():  # untouched scenarios keep their state bit for bit
        np.testing.assert_array_equal(tn(getattr(be, k))[1::2], v[1::2], err_msg=k)
    be.reset()
    be.rollout(30)
    for k, v in snap.items():  # same inputs -> same bits (no atomics on floats, no order dependence)
        np.testing.assert_array_equal(tn(getattr(be, k)), v, err_msg=k)
    be.close()


def test_ragged_and_empty_inputs(backend):
    """ragged agent counts, an agent-less scenario, an edge-less scenario, single-state trajectories."""
    from lcsim_b200 import packer, synth

    raws = [synth.make_raw_scenario(s, n) for s, n in ((11, 5), (12, 1), (13, 40))]
    scs = [packer.scenario_from_raw(r) for r in raws]
    scs[1].edge_xy = np.zeros((0, 2)); scs[1].edge_dir = np.zeros((0, 2))  # no road edges at all
    for sc in scs[:1]:  # single-state trajectories finish after their first update
        sc.traj_xy[0] = sc.traj_xy[0][:1]; sc.traj_vel[0] = sc.traj_vel[0][:1]; sc.traj_h[0] = sc.traj_h[0][:1]
    sb = packer.pack(scs)
    for reactive in (False, True):
        be = BatchEngine(sb, reactive=reactive, with_metrics=True, backend=backend)
        o = orc.Oracle(sb, reactive=reactive)
        for k in range(95):  # past total_steps on purpose (engine.py never stops the clock)
            be.step()
            o.step()
            o.metrics()
            compare_engine_with_oracle(be, o, POSE_TOL)
        be.close()


def test_degenerate_edge_geometry_falls_back(backend):
    """road edge = one circle around the scene: every cell near the centre is (almost) equidistant from all points, the
    candidate lists would explode and are dropped for that scenario; its searches take the hinted hash-grid path"""
    from lcsim_b200 import packer, proto, synth

    raw = synth.make_raw_scenario(31, 48)
    ang = np.linspace(0, 2 * np.pi, 2600)
    raw["edges"] = [(np.stack([150.0 * np.cos(ang), 150.0 * np.sin(ang)], axis=1).astype(np.float32).astype(np.float64),
                     proto.ROAD_EDGE_TYPE_BOUNDARY)]
    sb = packer.pack([packer.scenario_from_raw(raw), packer.scenario_from_raw(synth.make_raw_scenario(32, 48))])
    be = BatchEngine(sb, reactive=True, with_metrics=True, backend=backend)
    o = orc.Oracle(sb, reactive=True)
    for k in range(40):
        be.step()
        o.step()
        o.metrics()
        compare_engine_with_oracle(be, o, POSE_TOL, near_duplicate_edges=sb.edge_xy)
    be.close()


def test_abi_rejects_bad_calls(backend, synth8):
    import ctypes as C

    lib = backend.lib
    d = native.BatchDesc()
    d.abi_version, d.S, d.A, d.T, d.E = native.ABI_VERSION + 1, 1, 32, 8, 32
    h = C.c_void_p()
    assert lib.lcsim_b200_create(C.byref(d), C.byref(h)) == -1
    d.abi_version, d.A = native.ABI_VERSION, 4096
    assert lib.lcsim_b200_create(C.byref(d), C.byref(h)) == -1
    d.A = 32
    assert lib.lcsim_b200_create(C.byref(d), C.byref(h)) == 0
    assert lib.lcsim_b200_step(h, None, None, None, 0, None) == -3  # step before upload_static
    assert b"upload_static" in lib.lcsim_b200_last_error(h)
    lib.lcsim_b200_destroy(h)
    bad = synth8.scenario_slice(0, 1)
    bad.traj_len = bad.traj_len.copy()
    bad.traj_len[0, 0] = bad.T + 5
    with pytest.raises(native.NativeLibraryError, match="traj_len"):
        BatchEngine(bad, backend=backend)


@pytest.mark.parametrize("name", ["synth1_ads_bicycle_replay", "synth1_ads_bicycle_reactive"])
def test_env_step_matches_reference_golden(name, backend):
    """lcsim_b200_env_step == BicycleSingleEnv.step (gym_warpper.py:78-87): normalised ADS action in from host memory,
    reward terms / terminated / truncated / route out to host memory, per tick, against the reference's own values."""
    c = Case(name)
    be = BatchEngine(c.static, reactive=c.reactive, with_metrics=True, backend=backend)
    coef = {"forward": 2, "collision": 5, "smooth": 1, "destination": 1}  # lcsim/config/example.yml:22-26
    out = np.zeros(1, native.ENV_OUT_DTYPE)
    tn = be.backend.to_numpy
    for k in range(c.n_steps):
        act = np.ascontiguousarray(c.ads_actions[k].reshape(1, 2), np.float64)
        be.env_step(act, out, coef=coef, sync=True)
        ref = c.ref
        np.testing.assert_array_equal(ref["cursor"][k + 1], tn(be.cursor)[0], err_msg=f"tick {k + 1}: cursor")
        np.testing.assert_allclose(tn(be.pose64)[0][:, 0], ref["x"][k + 1], rtol=0, atol=POSE_TOL)
        np.testing.assert_allclose(out["terms"][0], ref["rewards"][k + 1], rtol=0, atol=1e-9, equal_nan=True)
        assert bool(out["terminated"][0]) == bool(ref["terminated"][k + 1])
        fin = ref["status"][k + 1][c.ads] in (native.IDLE, native.FINISHED)
        assert bool(out["finished"][0]) == fin
        assert bool(out["truncated"][0]) == (k + 1 >= c.static.total_steps or fin)
        if not fin:
            np.testing.assert_allclose(out["route"][0], ref["route"][k + 1], rtol=0, atol=POSE_TOL)
        want = sum(coef[key] * ref["rewards"][k + 1][i] for i, key in enumerate(native.REWARD_KEYS) if key in coef)
        np.testing.assert_allclose(out["reward"][0], want, rtol=0, atol=1e-9)
    be.close()
