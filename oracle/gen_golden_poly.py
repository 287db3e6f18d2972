"""Freezes the reference's polyline centres (Junction.roadgraph_data["poly_center_points"]) for the two bundled
WOMD scenarios into tests/golden_poly/*.npz, together with the raw lane centre-lines they were computed from.
Needs /root/reference (build container only):  python -m oracle.gen_golden_poly

The reference only builds the roadgraph when a MotionDiff model is loaded (engine.py:102-125); a stand-in whose
map_encoder returns zeros is injected, exactly as SURVEY.md §8c item 4 describes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcsim_b200 import packer, proto  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden_poly")


class _FakeEncoder:
    def forward(self, data):
        return torch.zeros(len(data["roadgraph"]["poly_center_points"]), 128)


class _FakeDiffuser:
    map_encoder = _FakeEncoder()

    def parameters(self):
        yield torch.zeros(1)


def main():
    Engine, _ = rh.load_reference()
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
        cp = junc.roadgraph_data["poly_center_points"].numpy()  # (P,7) float32
        sc = packer.scenario_from_pb(proto.load_map(os.path.join(d, "map.pb")), proto.load_agents(os.path.join(d, "agents.pb")))
        lanes = np.concatenate([p for p, _ in sc.lanes])
        lane_len = np.asarray([len(p) for p, _ in sc.lanes], np.int32)
        np.savez_compressed(os.path.join(OUT, f"poly_{scen[:6]}.npz"), lanes=lanes, lane_len=lane_len, centre=cp)
        print(scen, cp.shape)


if __name__ == "__main__":
    main()
