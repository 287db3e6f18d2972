"""Freezes encoder-input vectors into tests/golden_encoder/*.npz.
Needs /root/reference (build container only):  python -m oracle.gen_golden_encoder

For the two bundled WOMD scenarios the REAL engine is stepped 30 ticks; the history tensors of its active agents are
taken exactly as Junction.get_hetero_data builds them (junction/junction.py:301-334: float32 xyz / vel / heading /
shape / valid), the polyline centres from Junction.roadgraph_data (stand-in map encoder).  On those
tensors the input stage of AgentEncoder.forward (motion_diff/modules/agent_encoder.py:98-256) is evaluated with torch,
line for line, calling the reference's own angle_between_2d_vectors / wrap_angle (motion_diff/modules/utils.py,
loaded from its file).  The three third-party calls are restated from their call-site contract (parity unpinned at
that boundary: torch_cluster / torch_geometric are neither vendored nor installed, `pip download torch_cluster
--no-index --find-links /opt/wheelhouse` finds nothing):
  * torch_cluster.radius(x, y, r, max_num_neighbors) -> for every y the x with |x - y|^2 < r^2, ascending x, [y; x]
  * torch_cluster.radius_graph(x, r, loop=False)      -> for every target the sources, ascending, [source; target]
  * torch_geometric.utils.dense_to_sparse / subgraph  -> nonzero of the mask in row-major order / keep both-valid edges
Agent rows are in active-list order (the reference iterates a Python set; the graph is permutation equivariant)."""
import importlib.util
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as rh  # noqa: E402
from oracle.gen_golden_poly import _FakeDiffuser  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden_encoder")
H, TIME_SPAN, R = 11, 10, 50.0


def _utils():
    spec = importlib.util.spec_from_file_location("refmd_utils", os.path.join(rh.REFERENCE_ROOT, "motion_diff", "modules", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def radius(x, y, r, cap=300):
    rows = []
    for q in range(len(y)):
        d2 = ((x - y[q]) ** 2).sum(-1)
        idx = torch.nonzero(d2 < r * r).flatten()[:cap]
        rows.append(torch.stack([torch.full_like(idx, q), idx]))
    return torch.cat(rows, dim=1) if rows else torch.zeros(2, 0, dtype=torch.long)


def radius_graph(x, r, cap=300):
    rows = []
    for i in range(len(x)):
        d2 = ((x - x[i]) ** 2).sum(-1)
        idx = torch.nonzero(d2 < r * r).flatten()
        idx = idx[idx != i][:cap]
        rows.append(torch.stack([idx, torch.full_like(idx, i)]))
    return torch.cat(rows, dim=1) if rows else torch.zeros(2, 0, dtype=torch.long)


def main():
    Engine, _ = rh.load_reference()
    U = _utils()
    os.makedirs(OUT, exist_ok=True)
    for scen, ticks in (("1a1b6e6bacdd8880", 30), ("1a0df891a09ab6c8", 7)):
        d = os.path.join(rh.REFERENCE_ROOT, "examples", "data", "waymo", scen)
        eng = Engine(rh.make_config(d))
        fake = _FakeDiffuser()
        eng.lane_manager.init_polyline_emb(fake)
        eng.road_manager.init_polyline_emb(fake)
        eng.junction_manager.init_polyline_emb(fake)
        eng.junction_manager.init_roadgraph_data()
        eng.reset()
        for _ in range(ticks):
            eng.step()
        am = eng.agent_manager
        ids = list(am.agents.keys())
        act = am.active_agents
        st = lambda name: torch.from_numpy(np.stack([getattr(a.history_states, name) for a in act])).to(torch.float32)  # noqa: E731
        xyz, vel, heading, shape = st("xyz"), st("vel"), st("heading"), st("shape")
        valid = torch.from_numpy(np.stack([a.history_states.valid for a in act])).to(torch.bool)
        junc = eng.junction_manager.get_unique_junction()
        pl = junc.roadgraph_data["poly_center_points"]
        # ---- agent_encoder.py:101-166
        mask = valid[:, :H].squeeze(-1).contiguous()
        pos_a = xyz[:, :H, :2]
        motion = torch.cat([pos_a.new_zeros(len(act), 1, 2), pos_a[:, 1:] - pos_a[:, :-1]], dim=1)
        head_a = heading[:, :H].squeeze(-1).contiguous()
        head_vec = torch.stack([head_a.cos(), head_a.sin()], dim=-1)
        pos_pl = pl[:, :2].contiguous()
        head_pl = torch.atan2(pl[:, 4], pl[:, 3]).contiguous()
        v = vel[:, :H].contiguous()
        x_a = torch.stack([torch.norm(motion, p=2, dim=-1), U.angle_between_2d_vectors(ctr_vector=head_vec, nbr_vector=motion),
                           torch.norm(v, p=2, dim=-1), U.angle_between_2d_vectors(ctr_vector=head_vec, nbr_vector=v),
                           shape[:, :H, 0], shape[:, :H, 1]], dim=-1)
        # ---- :174-199
        pos_t, head_t, head_vec_t = pos_a.reshape(-1, 2), head_a.reshape(-1), head_vec.reshape(-1, 2)
        mask_t = mask.unsqueeze(2) & mask.unsqueeze(1)
        mask_t[..., :-1] = False
        A = len(act)
        nz = mask_t.nonzero()  # (a, i, j) row-major == dense_to_sparse of the block-diagonal adjacency
        e_t = torch.stack([nz[:, 0] * H + nz[:, 1], nz[:, 0] * H + nz[:, 2]])
        e_t = e_t[:, e_t[1] > e_t[0]]
        e_t = e_t[:, e_t[1] - e_t[0] <= TIME_SPAN]
        rel = pos_t[e_t[1]] - pos_t[e_t[0]]
        r_t = torch.stack([torch.norm(rel, p=2, dim=-1), U.angle_between_2d_vectors(ctr_vector=head_vec_t[e_t[1]], nbr_vector=rel),
                           U.wrap_angle(head_t[e_t[1]] - head_t[e_t[0]]), (e_t[0] - e_t[1]).float()], dim=-1)
        # ---- :202-231
        pos_s, head_s, head_vec_s, mask_s = pos_a[:, -1], head_a[:, -1], head_vec[:, -1], mask[:, -1]
        e_pl = radius(pos_s, pos_pl, R)
        e_pl = e_pl[:, mask_s[e_pl[1]]]
        rel = pos_pl[e_pl[0]] - pos_s[e_pl[1]]
        r_pl = torch.stack([torch.norm(rel, p=2, dim=-1), U.angle_between_2d_vectors(ctr_vector=head_vec_s[e_pl[1]], nbr_vector=rel),
                            U.wrap_angle(head_pl[e_pl[0]] - head_s[e_pl[1]])], dim=-1)
        # ---- :233-255
        e_aa = radius_graph(pos_s, R)
        e_aa = e_aa[:, mask_s[e_aa[0]] & mask_s[e_aa[1]]]
        rel = pos_s[e_aa[0]] - pos_s[e_aa[1]]
        r_aa = torch.stack([torch.norm(rel, p=2, dim=-1), U.angle_between_2d_vectors(ctr_vector=head_vec_s[e_aa[1]], nbr_vector=rel),
                            U.wrap_angle(head_s[e_aa[0]] - head_s[e_aa[1]])], dim=-1)
        np.savez_compressed(
            os.path.join(OUT, f"encoder_{scen[:6]}.npz"), ticks=np.int32(ticks),
            active=np.asarray([ids.index(a.id) for a in act], np.int32), xyz=xyz.numpy(), vel=vel.numpy(), heading=heading.numpy(),
            valid=valid.numpy(), poly_centre=pl.numpy(), x_a=x_a.numpy(), e_t=e_t.numpy().astype(np.int32), r_t=r_t.numpy(),
            e_pl2a=e_pl.numpy().astype(np.int32), r_pl2a=r_pl.numpy(), e_a2a=e_aa.numpy().astype(np.int32), r_a2a=r_aa.numpy())
        print(scen, "agents", A, "polylines", len(pl), "edges t/pl2a/a2a", e_t.shape[1], e_pl.shape[1], e_aa.shape[1])


if __name__ == "__main__":
    main()
