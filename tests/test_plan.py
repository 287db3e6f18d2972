"""Guidance costs with analytic gradients and batched plan metrics against vectors recorded
from the REAL reference (oracle/gen_golden_plan.py: Repeller / NoOffRoad losses with torch.autograd gradients,
compute_signed_distance_to_nearest_road_edge_point, compute_overlap_rate, compute_offroad_rate) on the two bundled
WOMD scenarios; plus the road-edge roadgraph tensors the engine derives from the packed edges."""
import glob
import os

import numpy as np
import pytest

from lcsim_b200 import native, packer
from lcsim_b200.batch import BatchEngine

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(glob.glob(os.path.join(HERE, "golden_plan", "*.npz")))
WOMD = {"1a1b6e": "womd_1a1b6e_replay", "1a0df8": "womd_1a0df8_replay"}


@pytest.fixture(scope="module", params=["hostemu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "cuda":
        return native.TorchCudaBackend(0)
    from hostemu import NumpyEmuBackend

    return NumpyEmuBackend()


def _static_for(path):
    """the packed scenario of the golden case (tests/golden/womd_*_replay.npz holds the same StaticBatch), with the
    road-edge polyline ids rebuilt from the reference's own ids"""
    from golden_util import Case

    key = os.path.basename(path)[len("plan_"):-len(".npz")]
    sb = Case(WOMD[key]).static
    z = np.load(path)
    ids = z["rg_ids"].astype(np.int64)
    sb.edge_poly_id = np.zeros((1, sb.E), np.int32)
    sb.edge_poly_id[0, : len(ids)] = ids - ids.min()
    return sb, z


@pytest.mark.parametrize("path", CASES)
def test_road_edge_tensors_match_reference(path):
    """(float32) packed edges == the type-15|16 part of Junction.roadgraph_points of the real engine"""
    sb, z = _static_for(path)
    n = int(sb.n_edge[0])
    assert n == len(z["rg_xyz"])
    np.testing.assert_array_equal(sb.edge_xy[0, :n].astype(np.float32), z["rg_xyz"][:, :2])
    np.testing.assert_array_equal(sb.edge_dir[0, :n].astype(np.float32), z["rg_dir"][:, :2])
    assert set(np.unique(z["rg_type"]).tolist()) <= {15, 16}
    # packer.pack derives the same polyline ordinals from the polyline starts
    ids = z["rg_ids"].astype(np.int64)
    assert (np.diff(ids) >= 0).all() and (np.diff(ids) <= 1).all()


@pytest.mark.parametrize("path", CASES)
def test_guidance_costs_and_gradients(backend, path):
    sb, z = _static_for(path)
    be = BatchEngine(sb, backend=backend)
    tn = be.backend.to_numpy
    A, T = z["xy"].shape[0], z["xy"].shape[1]
    xy = np.zeros((1, be.A, T, 2), np.float32)
    xy[0, :A] = z["xy"]
    r = be.guidance_cost("repeller", xy)
    np.testing.assert_allclose(tn(r["loss"])[0], z["repeller_loss"], rtol=2e-5)
    g = tn(r["grad"])[0, :A]
    scale = np.abs(z["repeller_grad"]).max()
    np.testing.assert_allclose(g, z["repeller_grad"], rtol=0, atol=2e-5 * scale)
    assert (np.abs(z["repeller_grad"]) > 0).sum() > 100  # the case exercises the hinge
    r = be.guidance_cost("no_offroad", xy, with_sd=True)
    np.testing.assert_allclose(tn(r["sd"])[0, :A], z["signed_distance"], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(tn(r["loss"])[0], z["no_offroad_loss"], rtol=2e-5)
    scale = np.abs(z["no_offroad_grad"]).max()
    np.testing.assert_allclose(tn(r["grad"])[0, :A], z["no_offroad_grad"], rtol=0, atol=2e-5 * scale)
    assert (np.abs(z["no_offroad_grad"]) > 0).sum() > 100
    be.close()


@pytest.mark.parametrize("path", CASES)
def test_plan_metrics(backend, path):
    sb, z = _static_for(path)
    be = BatchEngine(sb, backend=backend)
    tn = be.backend.to_numpy
    A, T = z["xy"].shape[0], z["xy"].shape[1]
    traj5 = np.zeros((1, be.A, T, 5), np.float32)
    traj5[0, :A, :, :2] = z["xy"]
    traj5[0, :A, :, 2:4] = z["lw"][:, None, :]
    traj5[0, :A, :, 4] = z["yaw"]
    mask = np.zeros((1, be.A), np.uint8)
    mask[0, :A] = z["mask"]
    m = be.plan_metrics(traj5, overlap_mask=mask, offroad_mask=mask)  # data["agent"]["type"] is 1 for every engine agent
    assert int(tn(m["overlap_count"])[0]) == int(z["overlap_count"])
    assert int(tn(m["offroad_count"])[0]) == int(z["offroad_count"])
    np.testing.assert_allclose(tn(m["overlap_rate"])[0], z["overlap_rate"], rtol=1e-6)
    np.testing.assert_allclose(tn(m["offroad_rate"])[0], z["offroad_rate"], rtol=1e-6)
    be.close()


def test_plan_kernels_need_polyline_ids(backend):
    from lcsim_b200 import workload

    sb = workload.synthetic_batch(40, 2, 16, workers=1)
    assert sb.edge_poly_id is not None and sb.edge_poly_id.max() > 10  # packer.pack fills them
    sb.edge_poly_id = None
    be = BatchEngine(sb, backend=backend)
    xy = np.zeros((2, be.A, 8, 2), np.float32)
    be.guidance_cost("repeller", xy)
    with pytest.raises(native.NativeLibraryError, match="edge_poly_id"):
        be.guidance_cost("no_offroad", xy)
    be.close()
