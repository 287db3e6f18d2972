"""numpy restatement of the lane-centreline projections — TEST INFRASTRUCTURE.

PINNED: tests/golden_lane/*.npz holds outputs of the reference's own compute_lane_heading_diff
(motion_diff/metrics/roadgraph.py:91-157) and OnLane.loss (motion_diff/modules/guidance.py:100-133) run on the real
engine's Junction.roadgraph_points (oracle/gen_golden_lane.py); test_lane.py holds this restatement to them.
The first-index argmin / lane id has no counterpart in the reference (WOMD agents carry lane=None) and is defined
here: argmin of the same float32 squared distance, first index on ties."""
import numpy as np

F = np.float32


def wrap_f32(a):
    pi = F(np.pi)
    return (-pi + np.mod((a + pi).astype(F), F(2) * pi)).astype(F)  # geometry.wrap_angle / guidance.wrap_angle


def d2(q, pts):
    dx, dy = (q[0] - pts[:, 0]).astype(F), (q[1] - pts[:, 1]).astype(F)
    return (dx * dx + dy * dy).astype(F)  # roadgraph.py:150 torch.sum((a - b) ** 2, -1)


def project(q, xy_theta, ids, top_k=64):
    """q = (x, y, heading) float32.  Returns lane_pt, lane_id, lane_dist, onlane_dist, heading_diff."""
    q = np.asarray(q, F)
    if len(xy_theta) == 0:
        return -1, -1, F(np.inf), F(np.inf), F(-1)
    dd = d2(q, xy_theta)
    k = int(np.argmin(dd))
    dist = np.sqrt(dd).astype(F)  # guidance.py:121 torch.norm
    ang = np.abs(wrap_f32((q[2] - xy_theta[:, 2]).astype(F)))
    gate = (ang < F(0.2)) & (dist < F(5.0))  # guidance.py:125-128
    on = dist[gate].min() if gate.any() else F(np.inf)
    hd = F(-1)
    if 0 < top_k <= len(dd):
        order = np.lexsort((np.arange(len(dd)), dd))[:top_k]  # roadgraph.py:151 topk(-distances, K); ties by index
        hd = ang[order].min()
    return k, int(ids[k]), dist[k], F(on), F(hd)
