"""Observation builder: polyline centres pinned to the reference, neighbour gather and history stack
against the numpy restatement of the call-site contract (tests/obs_oracle.py, parity unpinned at the torch_cluster
boundary), on both backends."""
import glob
import os

import numpy as np
import pytest

import obs_oracle
from lcsim_b200 import native, packer, synth, workload
from lcsim_b200.batch import BatchEngine
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module", params=["hostemu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "cuda":
        return native.TorchCudaBackend(0)
    from hostemu import NumpyEmuBackend

    return NumpyEmuBackend()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden_poly", "*.npz"))))
def test_polyline_centres_match_reference(path):
    """packer.polyline_centres == Junction.roadgraph_data['poly_center_points'] of the real reference (oracle/gen_golden_poly.py)."""
    z = np.load(path)
    off = np.concatenate([[0], np.cumsum(z["lane_len"])])
    lanes = [(z["lanes"][off[i]:off[i + 1]], i) for i in range(len(z["lane_len"]))]
    c = packer.polyline_centres(lanes)
    ref = z["centre"]
    assert c.shape[0] == ref.shape[0]
    np.testing.assert_array_equal(c[:, :2].astype(np.float32), ref[:, :2])
    np.testing.assert_array_equal(c[:, 2].astype(np.float32), np.arctan2(ref[:, 4], ref[:, 3]))


@pytest.mark.parametrize("sort,KA,KP", [(False, 16, 48), (True, 16, 48), (True, 8, 160), (True, 64, 128)])
def test_neighbour_gather_and_history(backend, sort, KA, KP):
    raws = [synth.make_raw_scenario(200 + i, 128) for i in range(3)]
    scs = [packer.scenario_from_raw(r) for r in raws]
    sb = packer.pack(scs)
    centres = [packer.polyline_centres(sc.lanes) for sc in scs]
    be = BatchEngine(sb, reactive=True, with_metrics=True, backend=backend)
    be.upload_polylines(centres)
    o = orc.Oracle(sb, reactive=True)
    tn = be.backend.to_numpy
    # small caps so that both the cap and the top-k cut are exercised: k <= 128 takes the radix select + small sort,
    # k > 128 the sort of all candidates
    for tick in range(60):
        be.step()
        o.step()
        if tick % 6:
            continue
        ego = sb.ads_index.copy() if tick % 12 == 0 else np.asarray([o.active[s, tick % max(1, o.n_active[s])] for s in range(sb.S)], np.int32)
        out = {k: tn(v) for k, v in be.build_obs(k_agents=KA, k_polys=KP, radius=50.0, sort_by_distance=sort, ego=ego).items()}
        ring_last = o.ring[:, :, -1, :]  # newest history entry of every agent (x, y, vx, vy, h)
        for s in range(sb.S):
            e = int(ego[s])
            n = int(o.n_active[s])
            act = o.active[s, :n]
            np.testing.assert_array_equal(out["history_agent"][s, :n], act)
            assert (out["history_agent"][s, n:] == -1).all()
            np.testing.assert_array_equal(out["history"][s, :n, :, :5], o.ring[s, act].astype(np.float32))
            valid = (np.arange(11)[None, :] >= 11 - o.ring_valid[s, act][:, None]).astype(np.float32)
            np.testing.assert_array_equal(out["history"][s, :n, :, 5], valid)
            if o.status[s, e] != orc.ACTIVE:
                assert out["n_nbr_agent"][s] == 0 and out["n_nbr_poly"][s] == 0
                assert (out["nbr_agent"][s] == -1).all() and (out["nbr_poly"][s] == -1).all()
                continue
            exyh = ring_last[s, e][[0, 1, 4]].astype(np.float32)
            others = act[act != e]
            idx, feat, cnt = obs_oracle.gather(exyh, others, ring_last[s, others][:, [0, 1, 4]].astype(np.float32), 50.0, KA, sort)
            assert cnt == out["n_nbr_agent"][s]
            np.testing.assert_array_equal(out["nbr_agent"][s], idx)
            np.testing.assert_allclose(out["nbr_agent_feat"][s], feat, rtol=3e-7, atol=2e-6)  # float32 ulps (contraction, sqrt/atan2/cos/sin)
            c = centres[s].astype(np.float32)
            idx, feat, cnt = obs_oracle.gather(exyh, np.arange(len(c), dtype=np.int32), c, 50.0, KP, sort)
            assert cnt == out["n_nbr_poly"][s]
            np.testing.assert_array_equal(out["nbr_poly"][s], idx)
            np.testing.assert_allclose(out["nbr_poly_feat"][s], feat, rtol=3e-7, atol=2e-6)
    be.close()


def test_build_obs_argument_errors(backend):
    sb = workload.synthetic_batch(300, 1, 16, workers=1)
    be = BatchEngine(sb, backend=backend)
    with pytest.raises(native.NativeLibraryError, match="upload_polylines"):
        be.build_obs()
    be.build_obs(with_polys=False)
    with pytest.raises(native.NativeLibraryError, match="k_agents"):
        be.build_obs(k_agents=5000, with_polys=False)
    be.close()


def test_env_step_builds_registered_observation(backend):
    """set_env_obs: every env_step leaves the observation of the NEW state in the registered tensors — the same bytes a
    separate build_obs writes after that tick; out= reuses the tensors of an earlier build_obs call."""
    raws = [synth.make_raw_scenario(260 + i, 64) for i in range(3)]
    scs = [packer.scenario_from_raw(r) for r in raws]
    sb = packer.pack(scs)
    centres = [packer.polyline_centres(sc.lanes) for sc in scs]
    be = BatchEngine(sb, reactive=False, with_metrics=True, backend=backend)
    be.upload_polylines(centres)
    tn = be.backend.to_numpy
    kw = dict(k_agents=8, k_polys=24, radius=50.0, sort_by_distance=True)
    # a second engine without a registered observation: its env_step is the fused single-tick launch, while the first
    # one's is split into tick items | metric items with the observation build beside the latter
    os.environ["LCSIM_B200_ENV_SPLIT"] = "1"  # (read at create; the default is the fused launch)
    try:
        be_split = BatchEngine(sb, reactive=False, with_metrics=True, backend=backend)
    finally:
        del os.environ["LCSIM_B200_ENV_SPLIT"]
    be_split.upload_polylines(centres)
    be.close()
    be = be_split
    be2 = BatchEngine(sb, reactive=False, with_metrics=True, backend=backend)
    reg = be.set_env_obs(**kw)
    out, out2 = np.zeros(sb.S, native.ENV_OUT_DTYPE), np.zeros(sb.S, native.ENV_OUT_DTYPE)
    rng = np.random.default_rng(5)
    sep = None
    for tick in range(25):
        act = rng.uniform(-0.5, 0.5, (sb.S, 2))
        be.env_step(act, out, coef=dict(forward=0.1, collision=10, offroad=5), sync=True)
        be2.env_step(act, out2, coef=dict(forward=0.1, collision=10, offroad=5), sync=True)
        assert out.tobytes() == out2.tobytes()
        for nm in ("pose64", "status", "coll_count", "nearest_edge", "offroad", "active", "stats"):
            np.testing.assert_array_equal(tn(getattr(be, nm)), tn(getattr(be2, nm)), err_msg=f"{nm} at tick {tick}")
        got = {k: tn(v).copy() for k, v in reg.items()}
        sep = be.build_obs(out=sep, **kw)
        be.sync()
        for k, v in sep.items():
            np.testing.assert_array_equal(got[k], tn(v), err_msg=f"{k} at tick {tick}")
    assert (got["n_nbr_poly"] > 0).any()
    be.set_env_obs(enable=False)
    before = {k: tn(v).copy() for k, v in reg.items()}
    be.env_step(None, out, sync=True)
    for k, v in reg.items():
        np.testing.assert_array_equal(before[k], tn(v))  # no longer written
    be.close()
    be2.close()


def test_env_step_device_equals_host_form(backend):
    """lcsim_b200_env_step_device: same records as the host-buffer form, written to a device buffer."""
    sb = workload.synthetic_batch(310, 3, 48, workers=1)
    a, b = BatchEngine(sb, backend=backend), BatchEngine(sb, backend=backend)
    tn = a.backend.to_numpy
    out = np.zeros(sb.S, native.ENV_OUT_DTYPE)
    rng = np.random.default_rng(9)
    dev = None
    for _ in range(15):
        act = rng.uniform(-1, 1, (sb.S, 2))
        a.env_step(act, out, coef=dict(forward=1.0, smooth=2.0), sync=True)
        dev = b.env_step_device(act, dev, coef=dict(forward=1.0, smooth=2.0))
        b.sync()
        got = np.frombuffer(tn(dev).tobytes(), native.ENV_OUT_DTYPE)
        for f in ("reward", "terms", "route", "terminated", "truncated", "finished"):  # (the pad bytes are not written)
            np.testing.assert_array_equal(got[f], out[f], err_msg=f)
    a.close()
    b.close()


@pytest.mark.gpu
def test_env_step_pinned_buffers_written_in_place():
    """Pinned host buffers are used in place (no staging copy): same records as with pageable buffers, for results only
    (LCSIM_B200_ENV_ZEROCOPY=1, the default) and for results + actions (=2)."""
    import torch

    sb = workload.synthetic_batch(330, 5, 64, workers=1)
    ref = BatchEngine(sb, device=0)
    out_ref = np.zeros(sb.S, native.ENV_OUT_DTYPE)
    engines = []
    for mode in ("1", "2"):
        os.environ["LCSIM_B200_ENV_ZEROCOPY"] = mode
        try:
            engines.append(BatchEngine(sb, device=0))
        finally:
            del os.environ["LCSIM_B200_ENV_ZEROCOPY"]
    pinned = [torch.zeros(sb.S * native.ENV_OUT_DTYPE.itemsize, dtype=torch.uint8).pin_memory() for _ in engines]
    act_pin = torch.zeros(sb.S, 2, dtype=torch.float64).pin_memory()
    rng = np.random.default_rng(3)
    coef = dict(forward=0.1, collision=10, offroad=5, smooth=2, destination=1)
    for _ in range(20):
        act = rng.uniform(-1, 1, (sb.S, 2))
        act_pin.copy_(torch.from_numpy(act))
        ref.env_step(act, out_ref, coef=coef, sync=True)
        for be, buf in zip(engines, pinned):
            be.env_step(act_pin, buf, coef=coef, sync=True)
            got = buf.numpy().view(native.ENV_OUT_DTYPE)
            for f in ("reward", "terms", "route", "terminated", "truncated", "finished"):
                np.testing.assert_array_equal(got[f], out_ref[f], err_msg=f)
    for be in engines + [ref]:
        be.close()
