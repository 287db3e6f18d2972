"""Freezes golden vectors from the REAL reference engine into tests/golden/*.npz.

Run in the build container only (needs /root/reference):  python -m oracle.gen_golden
Each file holds (i) the packed inputs (lcsim_b200.packer.StaticBatch arrays, so the case
can be replayed anywhere without the .pb files), (ii) the control flags and any external
inputs (ADS actions, re-plan tensors) and (iii) the reference's per-tick outputs recorded by
oracle/ref_harness.run_reference.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from lcsim_b200 import packer, synth  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
EXAMPLES = os.path.join(rh.REFERENCE_ROOT, "examples", "data", "waymo")

STATIC_FIELDS = ["n_agents", "ads_index", "agent_id", "agent_type", "length", "width", "dynamic", "traj_len",
                 "depart_tick", "wait_order", "home_xy", "traj_xy", "traj_vel", "traj_h", "n_edge", "edge_xy",
                 "edge_dir", "origin"]


def save_case(name, sb, ref, control, ads_actions=None, replan=None, n_steps=90):
    d = {"static_" + k: getattr(sb, k) for k in STATIC_FIELDS}
    d.update({"dims": np.asarray([sb.S, sb.A, sb.T, sb.E], np.int32),
              "clock": np.asarray([sb.start, sb.interval, sb.total_steps], np.float64),
              "control": np.asarray([int(control.get("reactive", False)), int(control.get("no_static", False)),
                                     int(control.get("allow_new_obj", True))], np.int32),
              "n_steps": np.asarray([n_steps], np.int32)})
    d.update({"ref_" + k: v for k, v in ref.items()})
    if ads_actions is not None:
        d["ads_actions"] = ads_actions
    if replan is not None:
        d["replan_tick"] = np.asarray([replan["tick"]], np.int32)
        for k in ("xy", "vel", "heading", "mask"):
            d["replan_" + k] = replan[k]
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, updates={int(ref['n_updates'].sum())}, "
          f"sum_active={int(ref['n_active'][1:].sum())} sum_coll={int(ref['coll'][1:].sum())} "
          f"sum_off={int(ref['offroad'][1:].sum())} sum_cursor={int(sum(ref['cursor'][k][ref['active'][k, :ref['n_active'][k]]].sum() for k in range(1, n_steps + 1)))}")


def main():
    assert rh.reference_available(), "needs /root/reference"
    # 1) the two bundled WOMD scenarios, both control modes (SURVEY.md §8c pins)
    for scen in ("1a1b6e6bacdd8880", "1a0df891a09ab6c8"):
        d = os.path.join(EXAMPLES, scen)
        sb = packer.pack_dirs([d])
        for reactive in (False, True):
            ref = rh.run_reference(d, 90, reactive=reactive)
            save_case(f"womd_{scen[:6]}_{'reactive' if reactive else 'replay'}", sb, ref, dict(reactive=reactive))
    # 2) synthetic WOMD-shaped scenarios through the same .pb bytes
    with tempfile.TemporaryDirectory() as tmp:
        def synth_case(seed, n_agents, name, ads_act=False, replan_tick=None, **control):
            d = os.path.join(tmp, f"s{seed}_{n_agents}")
            raw = synth.write_scenario(d, seed, n_agents)
            sb = packer.pack_dirs([d])
            sb_raw = packer.pack_raw([raw])
            for k in STATIC_FIELDS:  # the pb route and the raw route must pack identically
                assert np.array_equal(getattr(sb, k), getattr(sb_raw, k)), k
            rng = np.random.default_rng(1000 + seed)
            acts = rng.uniform(-1, 1, (90, 2)) if ads_act else None
            replan = None
            if replan_tick is not None:
                # stand-in for the denoiser output: fp32 (A,80,.) tensors that start near
                # each agent's logged track and move on smoothly
                A = sb.A
                base = sb.traj_xy[0][np.arange(A), np.minimum(replan_tick, sb.traj_len[0] - 1)]
                hd = sb.traj_h[0][np.arange(A), np.minimum(replan_tick, sb.traj_len[0] - 1)]
                sp = rng.uniform(0.0, 12.0, A)
                t = np.arange(80) * 0.1
                hh = hd[:, None] + rng.normal(0, 0.05, (A, 1)) * t[None, :]
                xy = base[:, None, :] + np.stack([np.cumsum(sp[:, None] * 0.1 * np.cos(hh), 1),
                                                  np.cumsum(sp[:, None] * 0.1 * np.sin(hh), 1)], -1)
                vel = np.stack([sp[:, None] * np.cos(hh), sp[:, None] * np.sin(hh)], -1)
                replan = dict(tick=replan_tick, xy=xy.astype(np.float32), vel=vel.astype(np.float32),
                              heading=hh.astype(np.float32), mask=(rng.random(A) < 0.7))
            ref = rh.run_reference(d, 90, ads_actions=acts, replan=replan, **control)
            assert np.array_equal(ref["dynamic"], sb.dynamic[0]), "dynamic_flag mismatch"
            save_case(name, sb, ref, control, acts, replan)

        synth_case(0, 128, "synth0_replay", reactive=False)
        synth_case(0, 128, "synth0_reactive", reactive=True)
        synth_case(1, 64, "synth1_ads_bicycle_replay", ads_act=True, reactive=False)
        synth_case(1, 64, "synth1_ads_bicycle_reactive", ads_act=True, reactive=True)
        synth_case(2, 96, "synth2_nostatic_reactive", reactive=True, no_static=True)
        synth_case(3, 96, "synth3_nonew_replay", reactive=False, allow_new_obj=False)
        synth_case(4, 64, "synth4_replan_replay", replan_tick=10, reactive=False)
        synth_case(4, 64, "synth4_replan_reactive", replan_tick=10, reactive=True)


if __name__ == "__main__":
    main()
