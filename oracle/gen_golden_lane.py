"""Freezes lane-centreline projection vectors from the REAL reference into tests/golden_lane/*.npz.
Needs /root/reference (build container only):  python -m oracle.gen_golden_lane

What is recorded, for the two bundled WOMD scenarios:
  * Junction.roadgraph_points {xyz, dir, type, ids} (junction/junction.py:186-253) built by the real engine with a
    stand-in map encoder -> pins lcsim_b200.packer.centreline_points;
  * per query (agent poses sampled from the scenario's own trajectories, float32 like data["agent"]):
      - compute_lane_heading_diff (motion_diff/metrics/roadgraph.py:91-157) called on ONE agent / ONE step, so the
        returned scalar is that query's min |wrap(dheading)| over its 64 nearest centreline points;
      - OnLane.loss (motion_diff/modules/guidance.py:100-133) with a stand-in diffuser whose _reconstruct_traj
        returns the queries; torch.mean is intercepted to keep the per-query min_dis vector.
    The two reference modules are loaded from their files (their package __init__ pulls in pytorch_lightning /
    torch_cluster, which are not installed); torch_geometric.data is the stub of oracle/ref_harness.py."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcsim_b200 import packer, proto  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from oracle.gen_golden_poly import _FakeDiffuser  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden_lane")


def _load_reference_modules():
    root = os.path.join(rh.REFERENCE_ROOT, "motion_diff")
    for name, sub in (("refmd", ""), ("refmd.metrics", "metrics"), ("refmd.modules", "modules")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(root, sub)]
        sys.modules[name] = m

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    geometry = load("refmd.metrics.geometry", "metrics/geometry.py")
    sys.modules["refmd.metrics"].geometry = geometry
    roadgraph = load("refmd.metrics.roadgraph", "metrics/roadgraph.py")
    sys.modules["refmd.metrics"].roadgraph = roadgraph
    guidance = load("refmd.modules.guidance", "modules/guidance.py")
    return roadgraph, guidance


class _Data(dict):
    """HeteroData stand-in: nested dict access is all the two functions use."""


def main():
    Engine, _ = rh.load_reference()
    roadgraph, guidance = _load_reference_modules()
    os.makedirs(OUT, exist_ok=True)
    for scen in ("1a1b6e6bacdd8880", "1a0df891a09ab6c8"):
        d = os.path.join(rh.REFERENCE_ROOT, "examples", "data", "waymo", scen)
        eng = Engine(rh.make_config(d))
        fake = _FakeDiffuser()
        eng.lane_manager.init_polyline_emb(fake)
        eng.road_manager.init_polyline_emb(fake)
        eng.junction_manager.init_polyline_emb(fake)
        eng.junction_manager.init_roadgraph_data()
        junc = eng.junction_manager.get_unique_junction()
        rg = {k: v for k, v in junc.roadgraph_points.items() if k != "num_nodes"}
        sc = packer.scenario_from_pb(proto.load_map(os.path.join(d, "map.pb")), proto.load_agents(os.path.join(d, "agents.pb")))
        # queries: every agent's reference trajectory at three of its states, float32 (data["agent"] tensors are float32)
        q = []
        for a in range(len(sc.traj_xy)):
            n = len(sc.traj_h[a])
            for k in sorted({0, n // 2, n - 1}):
                q.append([sc.traj_xy[a][k][0], sc.traj_xy[a][k][1], sc.traj_h[a][k]])
        q = np.asarray(q, np.float32)
        nq = len(q)
        # ---- compute_lane_heading_diff, one query per call
        hd = np.zeros(nq, np.float32)
        for i in range(nq):
            data = _Data(agent={"xyz": torch.zeros(1, 1, 3), "type": torch.ones(1, dtype=torch.int64)},
                         roadgraph_points={"xyz": rg["xyz"], "dir": rg["dir"], "type": rg["type"]})
            traj = torch.tensor([[[q[i, 0], q[i, 1], 4.5, 2.0, q[i, 2]]]], dtype=torch.float32)
            hd[i] = float(roadgraph.compute_lane_heading_diff(data, traj, torch.ones(1, dtype=torch.bool)))
        # ---- OnLane.loss: queries as (nq, 2, 2) tracks p -> p + (cos h, sin h) * 1e-3... the loss derives the query
        # direction from consecutive points, so each query becomes a two-point track along its heading
        step = np.stack([np.cos(q[:, 2]), np.sin(q[:, 2])], axis=1).astype(np.float32)
        xy = torch.from_numpy(np.stack([q[:, :2], q[:, :2] + step], axis=1))  # (nq, 2, 2)

        class _Diff:
            def _reconstruct_traj(self, data, x):
                return x

        captured = {}
        real_mean = torch.mean

        def spy(t, *a, **k):
            captured["min_dis"] = t.detach().clone()
            return real_mean(t, *a, **k)

        torch.mean = spy
        try:
            guidance.OnLane().loss(xy, _Diff(), _Data(roadgraph_points={"xyz": rg["xyz"], "dir": rg["dir"], "type": rg["type"]}))
        finally:
            torch.mean = real_mean
        dir_xy = torch.cat([xy[:, 1:, :] - xy[:, :-1, :], xy[:, -1:, :] - xy[:, -2:-1, :]], dim=1).reshape(-1, 2)
        on_q = np.concatenate([xy.reshape(-1, 2).numpy(), torch.atan2(dir_xy[:, 1], dir_xy[:, 0]).numpy()[:, None]], axis=1)
        lanes = np.concatenate([p for p, _ in sc.lanes])
        lane_len = np.asarray([len(p) for p, _ in sc.lanes], np.int32)
        lines = np.concatenate([p for p, _ in sc.lines]) if sc.lines else np.zeros((0, 2))
        line_len = np.asarray([len(p) for p, _ in sc.lines], np.int32)
        line_type = np.asarray([t for _, t in sc.lines], np.int32)
        m = (rg["type"] == 1) | (rg["type"] == 2)
        rg = {k: v[m] for k, v in rg.items()}  # only the centreline-typed points are searched; keep the fixture small
        np.savez_compressed(
            os.path.join(OUT, f"lane_{scen[:6]}.npz"), lanes=lanes, lane_len=lane_len, lines=lines, line_len=line_len, line_type=line_type,
            rg_xyz=rg["xyz"].numpy(), rg_dir=rg["dir"].numpy(), rg_type=rg["type"].numpy(), rg_ids=rg["ids"].numpy(),
            hd_query=q, hd_value=hd, on_query=on_q.astype(np.float32), on_value=captured["min_dis"].numpy())
        print(scen, "roadgraph points", len(rg["xyz"]), "queries", nq, "finite onlane", int(np.isfinite(captured["min_dis"].numpy()).sum()))


if __name__ == "__main__":
    main()
