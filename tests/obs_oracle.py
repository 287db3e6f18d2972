"""numpy restatement of the neighbour gather at the reference's call sites — TEST INFRASTRUCTURE.

parity unpinned: the reference delegates this arithmetic to torch_cluster==1.6.3 (requirements.txt:15;
motion_diff/modules/agent_encoder.py:208-215 `radius`, :235-241 `radius_graph`), which is neither vendored under
/root/reference nor installed here, and no reference test pins its output.  What is restated is the call-site
contract: float32 positions, strict d^2 < r^2, no self loop, ascending source order, at most `cap` per query, then
the reference's own relation features (agent_encoder.py:217-231,243-255; motion_diff/modules/utils.py:106-119).
The polyline centres themselves ARE pinned (tests/golden_poly, from Junction.init_roadgraph_data)."""
import numpy as np

F = np.float32


def wrap_f32(a):
    pi = F(np.pi)
    return (-pi + np.mod(a + pi, F(2) * pi)).astype(F)  # min_val + (angle + max_val) % (max_val - min_val)


def rel_feature(ex, ey, eh, nx, ny, nh):
    rx, ry = (nx - ex).astype(F), (ny - ey).astype(F)
    dist = np.sqrt(rx * rx + ry * ry).astype(F)
    hx, hy = np.cos(eh).astype(F), np.sin(eh).astype(F)
    ang = np.arctan2(hx * ry - hy * rx, (F(0) + hx * rx) + hy * ry).astype(F)  # torch .sum() accumulates from +0
    return np.stack([dist, ang, wrap_f32((nh - eh).astype(F))], axis=-1)


def gather(ego_xyh, ids, xyh, radius, cap, sort_by_distance):
    """ego_xyh: (3,) float32; ids (n,), xyh (n,3) float32 candidates in source order.  Returns (idx (cap,), feat (cap,3), n)."""
    idx = np.full(cap, -1, np.int32)
    feat = np.zeros((cap, 3), F)
    if len(ids) == 0:
        return idx, feat, 0
    dx, dy = (xyh[:, 0] - ego_xyh[0]).astype(F), (xyh[:, 1] - ego_xyh[1]).astype(F)
    d2 = (dx * dx + dy * dy).astype(F)
    sel = np.nonzero(d2 < F(radius) * F(radius))[0]
    if sort_by_distance:
        sel = sel[np.lexsort((sel, d2[sel]))]
    n = len(sel)
    sel = sel[:cap]
    idx[: len(sel)] = ids[sel]
    feat[: len(sel)] = rel_feature(ego_xyh[0], ego_xyh[1], ego_xyh[2], xyh[sel, 0], xyh[sel, 1], xyh[sel, 2])
    return idx, feat, n
