"""Freezes guidance-cost and plan-metric vectors from the REAL reference into tests/golden_plan/*.npz.
Needs /root/reference (build container only):  python -m oracle.gen_golden_plan

For the two bundled WOMD scenarios: the road-edge part of Junction.roadgraph_points (xyz, dir, ids, type 15|16) built
by the real engine (pins what lcsim_b200_upload_static derives from the packed edges), random float32 plans (A, 80, 2)
that start on the agents' logged tracks, and on them
  * Repeller().loss / NoOffRoad().loss (motion_diff/modules/guidance.py:17-85) with torch.autograd gradients,
  * compute_signed_distance_to_nearest_road_edge_point (motion_diff/metrics/roadgraph.py:160-232),
  * compute_overlap_rate (metrics/overlap.py:7-43) and compute_offroad_rate (metrics/roadgraph.py:36-88).
The reference modules are loaded from their files like oracle/gen_golden_lane.py does."""
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

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden_plan")
T_PLAN = 80


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
    overlap = load("refmd.metrics.overlap", "metrics/overlap.py")
    guidance = load("refmd.modules.guidance", "modules/guidance.py")
    return roadgraph, overlap, guidance


class _Data(dict):
    """HeteroData stand-in: nested dict access is all these functions use."""


class _Diff:
    def _reconstruct_traj(self, data, x):
        return x


def main():
    Engine, _ = rh.load_reference()
    roadgraph, overlap, guidance = _load_reference_modules()
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
        A = len(sc.traj_xy)
        rng = np.random.default_rng(int(scen[:6], 16) % 100000)
        # plans: from a state of the logged track, constant speed along a slowly turning heading; a third of the agents
        # are started 1-3 m from another agent so that the repeller and the overlap test have work
        start = np.stack([sc.traj_xy[a][rng.integers(0, len(sc.traj_xy[a]))] for a in range(A)])
        head = np.asarray([sc.traj_h[a][0] for a in range(A)])
        for a in range(0, A, 3):
            b = int(rng.integers(0, A))
            if b != a:
                start[a] = start[b] + rng.uniform(-3, 3, 2)
        speed = rng.uniform(0.0, 10.0, A)
        hh = head[:, None] + np.cumsum(rng.normal(0, 0.01, (A, T_PLAN)), axis=1)
        step = np.stack([np.cos(hh), np.sin(hh)], -1) * speed[:, None, None] * 0.1
        xy = (start[:, None, :] + np.cumsum(step, axis=1)).astype(np.float32)
        yaw = hh.astype(np.float32)
        lw = np.stack([sc.length, sc.width], axis=1).astype(np.float32)
        data = _Data(agent={"xyz": torch.zeros(A, 1, 3), "type": torch.ones(A, dtype=torch.int64),
                            "shape": torch.from_numpy(np.concatenate([lw, np.ones((A, 1), np.float32)], 1))[:, None, :]},
                     roadgraph_points={"xyz": rg["xyz"], "dir": rg["dir"], "type": rg["type"], "ids": rg["ids"]})
        out = {}
        for name, mod in (("repeller", guidance.Repeller()), ("no_offroad", guidance.NoOffRoad())):
            x = torch.from_numpy(xy).clone().requires_grad_(True)
            loss = mod.loss(x, _Diff(), data)
            (g,) = torch.autograd.grad(loss, x)
            out[name + "_loss"] = np.float32(loss.item())
            out[name + "_grad"] = g.numpy()
        pts = torch.cat([rg["xyz"], rg["dir"], rg["ids"].unsqueeze(-1), rg["type"].unsqueeze(-1)], dim=-1).contiguous()
        out["signed_distance"] = roadgraph.compute_signed_distance_to_nearest_road_edge_point(
            torch.from_numpy(xy).reshape(-1, 2), pts).numpy().reshape(A, T_PLAN)
        traj5 = torch.from_numpy(np.concatenate([xy, np.repeat(lw[:, None, :], T_PLAN, 1), yaw[..., None]], axis=-1))
        mask = torch.from_numpy(rng.random(A) < 0.8)
        out["overlap_rate"] = np.float32(overlap.compute_overlap_rate(data, traj5, mask).item())
        out["offroad_rate"] = np.float32(roadgraph.compute_offroad_rate(data, traj5, mask).item())
        cnt = 0
        for i in range(T_PLAN):
            cnt += int(torch.sum(overlap.compute_pairwise_overlaps(traj5[:, i]) & mask))
        out["overlap_count"] = np.int32(cnt)
        out["offroad_count"] = np.int32(int((torch.from_numpy(out["signed_distance"])[mask] > 0).sum()))
        m = (rg["type"] == 15) | (rg["type"] == 16)
        np.savez_compressed(os.path.join(OUT, f"plan_{scen[:6]}.npz"), xy=xy, yaw=yaw, lw=lw, mask=mask.numpy(),
                            rg_xyz=rg["xyz"][m].numpy(), rg_dir=rg["dir"][m].numpy(), rg_ids=rg["ids"][m].numpy(),
                            rg_type=rg["type"][m].numpy(), **out)
        print(scen, "agents", A, "edge points", int(m.sum()), {k: (float(v) if np.ndim(v) == 0 else v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
