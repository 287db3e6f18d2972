"""Lane-centreline projection: centreline points pinned to the reference's Junction.roadgraph_points,
the numpy restatement pinned to outputs of the reference's compute_lane_heading_diff / OnLane.loss
(tests/golden_lane, oracle/gen_golden_lane.py), and lcsim_b200_project_lanes against the restatement on both
backends: integer outputs (point index, lane id) bit-exact, float32 outputs equal."""
import glob
import os

import numpy as np
import pytest

import lane_oracle
from golden_util import Case
from lcsim_b200 import native, packer, synth
from lcsim_b200.batch import BatchEngine

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(glob.glob(os.path.join(HERE, "golden_lane", "*.npz")))


@pytest.fixture(scope="module", params=["hostemu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "cuda":
        return native.TorchCudaBackend(0)
    from hostemu import NumpyEmuBackend

    return NumpyEmuBackend()


def _points(z):
    off = np.concatenate([[0], np.cumsum(z["lane_len"])])
    lanes = [(z["lanes"][off[i]:off[i + 1]], i) for i in range(len(z["lane_len"]))]
    off = np.concatenate([[0], np.cumsum(z["line_len"])])
    lines = [(z["lines"][off[i]:off[i + 1]], int(z["line_type"][i])) for i in range(len(z["line_len"]))]
    return packer.centreline_points(lanes, lines)


@pytest.mark.parametrize("path", GOLD)
def test_centreline_points_match_reference(path):
    z = np.load(path)
    xyt, ids, xy_dir = _points(z)  # the fixture holds the reference's points of type 1|2
    assert len(xyt) == len(z["rg_xyz"]) and set(z["rg_type"].tolist()) <= {1, 2}
    np.testing.assert_array_equal(xy_dir[:, :2], z["rg_xyz"][:, :2])
    np.testing.assert_array_equal(xy_dir[:, 2:], z["rg_dir"][:, :2])
    np.testing.assert_array_equal(ids, z["rg_ids"])


@pytest.mark.parametrize("path", GOLD)
def test_restatement_matches_reference_functions(path):
    z = np.load(path)
    xyt, ids, _ = _points(z)
    for q, want in zip(z["hd_query"], z["hd_value"]):
        got = lane_oracle.project(q, xyt, ids)[4]
        np.testing.assert_allclose(got, want, rtol=0, atol=5e-7)  # float32 atan2 of torch vs numpy: <= 1 ulp of pi
    n_fin = 0
    for q, want in zip(z["on_query"], z["on_value"]):
        got = lane_oracle.project(q, xyt, ids)[3]
        if np.isfinite(want):
            n_fin += 1
            # torch.norm (guidance.py:121) may contract x*x + y*y into an fma; 2 float32 ulps of a <5 m distance
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)
        else:
            assert not np.isfinite(got)
    assert n_fin > 50


def _check_engine(be, points, tn, top_k):
    out = {k: tn(v) for k, v in be.project_lanes(top_k=top_k).items()}
    status, pose = tn(be.status), tn(be.pose64)
    n_checked = 0
    for s in range(be.S):
        xyt, ids = points[s]
        for a in range(be.A):
            if status[s, a] != native.ACTIVE:
                assert out["lane_pt"][s, a] == -1 and out["lane_id"][s, a] == -1 and out["heading_diff"][s, a] == -1
                assert np.isinf(out["lane_dist"][s, a]) and np.isinf(out["onlane_dist"][s, a])
                continue
            k, lid, dist, on, hd = lane_oracle.project(pose[s, a, :3].astype(np.float32), xyt, ids, top_k)
            assert out["lane_pt"][s, a] == k and out["lane_id"][s, a] == lid, (s, a)
            assert out["lane_dist"][s, a] == dist and out["onlane_dist"][s, a] == on, (s, a)
            assert out["heading_diff"][s, a] == hd, (s, a, out["heading_diff"][s, a], hd)
            n_checked += 1
    return n_checked


def test_project_lanes_on_reference_scenario(backend):
    """the bundled WOMD scenario, its real lane centrelines (14 k points), every active agent at several ticks"""
    c = Case("womd_1a1b6e_replay")
    z = np.load(os.path.join(HERE, "golden_lane", "lane_1a1b6e.npz"))
    xyt, ids, _ = _points(z)
    be = BatchEngine(c.static, reactive=False, with_metrics=True, backend=backend)
    be.upload_centrelines([(xyt, ids)])
    n = 0
    for k in range(3):
        n += _check_engine(be, [(xyt, ids)], be.backend.to_numpy, 64)
        be.rollout(30)
    assert n > 60
    be.close()


def test_project_lanes_synthetic_and_edge_cases(backend):
    """ragged scenarios: one with fewer centreline points than top_k (heading_diff = -1, roadgraph.py:144-145),
    one without any, one full synthetic map; smaller top_k."""
    raws = [synth.make_raw_scenario(300 + i, n) for i, n in enumerate((40, 7, 25))]
    scs = [packer.scenario_from_raw(r) for r in raws]
    sb = packer.pack(scs)
    pts = [packer.centreline_points(sc.lanes, sc.lines)[:2] for sc in scs]
    pts[1] = (pts[1][0][:20], pts[1][1][:20])
    pts[2] = (pts[2][0][:0], pts[2][1][:0])
    be = BatchEngine(sb, reactive=True, with_metrics=True, backend=backend)
    be.upload_centrelines(pts)
    be.rollout(12)
    assert _check_engine(be, pts, be.backend.to_numpy, 64) > 20
    assert _check_engine(be, pts, be.backend.to_numpy, 16) > 20
    be.close()
