"""Encoder inputs for all listed agents against vectors built from the REAL engine's history tensors with
the input stage of AgentEncoder.forward evaluated in torch (oracle/gen_golden_encoder.py: the reference's own
angle_between_2d_vectors / wrap_angle; torch_cluster's radius calls restated from their call-site contract).  Also
pins the (N,11,.) history stack of lcsim_b200_build_obs to the reference's float32 history tensors."""
import glob
import os

import numpy as np
import pytest

from golden_util import Case
from lcsim_b200 import native
from lcsim_b200.batch import BatchEngine

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(glob.glob(os.path.join(HERE, "golden_encoder", "*.npz")))
WOMD = {"1a1b6e": "womd_1a1b6e_replay", "1a0df8": "womd_1a0df8_replay"}
# float32 ulps: contraction of the 2-term dot products, sqrt / atan2 / cos / sin implementations (torch vs CUDA / libm)
FEAT_TOL = dict(rtol=3e-6, atol=3e-6)


@pytest.fixture(scope="module", params=["hostemu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "cuda":
        return native.TorchCudaBackend(0)
    from hostemu import NumpyEmuBackend

    return NumpyEmuBackend()


@pytest.mark.parametrize("path", CASES)
def test_encoder_inputs_match_reference(backend, path):
    z = np.load(path)
    sb = Case(WOMD[os.path.basename(path)[len("encoder_"):-len(".npz")]]).static
    be = BatchEngine(sb, backend=backend)
    centres = z["poly_centre"]
    be.upload_polylines([np.stack([centres[:, 0], centres[:, 1], np.arctan2(centres[:, 4], centres[:, 3])], axis=1)])
    for _ in range(int(z["ticks"])):
        be.step()
    tn = be.backend.to_numpy
    n = len(z["active"])
    # the history stack of build_obs == the float32 tensors Junction.get_hetero_data builds, agent for agent
    obs = {k: tn(v) for k, v in be.build_obs(with_polys=False).items()}
    np.testing.assert_array_equal(obs["history_agent"][0, :n], z["active"])
    hist = obs["history"][0, :n]
    np.testing.assert_array_equal(hist[..., 0:2], z["xyz"][..., :2])
    np.testing.assert_array_equal(hist[..., 2:4], z["vel"])
    np.testing.assert_array_equal(hist[..., 4], z["heading"][..., 0])
    np.testing.assert_array_equal(hist[..., 5] > 0, z["valid"][..., 0])
    out = {k: tn(v) for k, v in be.build_encoder_inputs().items()}
    np.testing.assert_array_equal(out["agent"][0, :n], z["active"])
    assert (out["agent"][0, n:] == -1).all()
    np.testing.assert_allclose(out["x_a"][0, :n], z["x_a"], **FEAT_TOL)
    # temporal edges: (row, i) -> (row, 10) in row-major order
    H = native.H_STEPS
    rows, steps = np.nonzero(out["r_t_valid"][0, :n])
    e_t = np.stack([rows * H + steps, rows * H + (H - 1)])
    np.testing.assert_array_equal(e_t, z["e_t"])
    np.testing.assert_allclose(out["r_t"][0, :n][rows, steps], z["r_t"], **FEAT_TOL)
    assert not out["r_t_valid"][0, n:].any()
    # a2a: CSR over targets -> [source; target] sorted by target then source
    off = out["a2a_offset"][0]
    assert off[0] == 0 and off[-1] == out["n_a2a"][0] == z["e_a2a"].shape[1]
    tgt = np.repeat(np.arange(len(off) - 1), np.diff(off))
    np.testing.assert_array_equal(np.stack([out["a2a_src"][0, : off[-1]], tgt]), z["e_a2a"])
    np.testing.assert_allclose(out["a2a_feat"][0, : off[-1]], z["r_a2a"], **FEAT_TOL)
    # pl2a: CSR over polylines -> [polyline; agent]
    off = out["pl2a_offset"][0]
    assert off[-1] == out["n_pl2a"][0] == z["e_pl2a"].shape[1]
    pl = np.repeat(np.arange(len(off) - 1), np.diff(off))
    np.testing.assert_array_equal(np.stack([pl, out["pl2a_agent"][0, : off[-1]]]), z["e_pl2a"])
    np.testing.assert_allclose(out["pl2a_feat"][0, : off[-1]], z["r_pl2a"], **FEAT_TOL)
    be.close()


def test_encoder_inputs_capacity_and_batch(backend):
    """edges beyond a capacity are counted but not stored; several scenarios at once; scenarios without agents"""
    from lcsim_b200 import workload

    sb, centres = workload.synthetic_batch(70, 3, 48, workers=1, with_centres=True)
    be = BatchEngine(sb, backend=backend)
    be.upload_polylines([c[:5] for c in centres])  # a second upload replaces the first (and frees its buffers)
    be.upload_polylines(centres)
    be.rollout(12)
    tn = be.backend.to_numpy
    full = {k: tn(v) for k, v in be.build_encoder_inputs(cap_a2a=48 * 48, cap_pl2a=60000).items()}
    small = {k: tn(v) for k, v in be.build_encoder_inputs(cap_a2a=10, cap_pl2a=7).items()}
    np.testing.assert_array_equal(full["n_a2a"], small["n_a2a"])
    np.testing.assert_array_equal(full["n_pl2a"], small["n_pl2a"])
    np.testing.assert_array_equal(full["a2a_offset"], small["a2a_offset"])
    np.testing.assert_array_equal(full["a2a_src"][:, :10], small["a2a_src"])
    np.testing.assert_array_equal(full["pl2a_agent"][:, :7], small["pl2a_agent"])
    assert (full["n_a2a"] > 10).all() and (full["n_pl2a"] > 100).all()
    # symmetric radius graph: every edge has its reverse
    for s in range(sb.S):
        off = full["a2a_offset"][s]
        tgt = np.repeat(np.arange(len(off) - 1), np.diff(off))
        pairs = set(zip(full["a2a_src"][s, : off[-1]].tolist(), tgt.tolist()))
        assert all((b, a) in pairs for a, b in pairs)
    be.close()
