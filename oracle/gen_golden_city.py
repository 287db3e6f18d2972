"""Freezes LaneIDM / MOSS-shaped golden vectors from the REAL reference into tests/golden_city/*.npz.
Needs /root/reference (build container only):  python -m oracle.gen_golden_city

Each case: lcsim_b200.city.make_city -> map.pb / agents.pb in a temp dir -> reference Engine(config).reset() ->
n x Engine.step(); after reset and after every step it records the active list, and for every agent status, lane pb
id, s, v, xy, heading, the LaneIDM observation (ahead_vehicle), check_collision_all() counts and Agent.is_offroad().
The packed static arrays travel in the same file, so the tests do not depend on the generator staying unchanged."""
import dataclasses
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcsim_b200 import city  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden_city")

CASES = {
    "city_g3_a60": dict(seed=1, grid=3, n_agents=60, n_ticks=260),
    "city_g2_a24": dict(seed=2, grid=2, n_agents=24, n_ticks=200, max_hops=5),
    "city_g4_a150_dense": dict(seed=3, grid=4, n_agents=150, n_ticks=180, depart_window=5.0),
    # movements that only some lanes offer + random end lanes: forced lane changes (idm_policy.py:118-167)
    "city_g3_a80_lc": dict(seed=5, grid=3, n_agents=80, n_ticks=280, require_lc=True, max_hops=5),
    "city_g4_a120_lc": dict(seed=7, grid=4, n_agents=120, n_ticks=240, require_lc=True, max_hops=5, depart_window=8.0),
}


def run_reference_city(data_dir: str, n_ticks: int, total: int = 100000) -> dict:
    import contextlib
    import io

    Engine, _ = rh.load_reference()
    cfg = rh.make_config(data_dir)
    cfg["control"]["step"]["total"] = total
    with contextlib.redirect_stdout(io.StringIO()):
        eng = Engine(cfg)
    eng.reset()
    am = eng.agent_manager
    ids = list(am.agents.keys())
    index_of = {aid: i for i, aid in enumerate(ids)}
    A, K = len(ids), n_ticks
    out = dict(
        n_active=np.zeros(K + 1, np.int32), active=np.full((K + 1, A), -1, np.int32), status=np.zeros((K + 1, A), np.int8),
        lane=np.zeros((K + 1, A), np.int32), s=np.zeros((K + 1, A)), v=np.zeros((K + 1, A)), x=np.zeros((K + 1, A)),
        y=np.zeros((K + 1, A)), h=np.zeros((K + 1, A)), lead_valid=np.zeros((K + 1, A), np.uint8),
        lead_dis=np.zeros((K + 1, A)), lead_v=np.zeros((K + 1, A)), coll=np.zeros((K + 1, A), np.int32),
        offroad=np.zeros((K + 1, A), np.uint8), is_lc=np.zeros((K + 1, A), np.uint8))

    def record(k):
        act = am.active_agents
        out["n_active"][k] = len(act)
        for i, a in enumerate(act):
            out["active"][k, i] = index_of[a.id]
        for aid, a in am.agents.items():
            i = index_of[aid]
            out["status"][k, i] = a.runtime.status.value
            out["lane"][k, i] = a.runtime.lane.id
            out["s"][k, i], out["v"][k, i] = a.runtime.s, a.runtime.v
            out["x"][k, i], out["y"][k, i], out["h"][k, i] = a.runtime.xy[0], a.runtime.xy[1], a.runtime.heading
            out["is_lc"][k, i] = bool(a.policy.lc_runtime.is_lc)
        if len(act):
            cc = am.check_collision_all()
            for i, a in enumerate(act):
                j = index_of[a.id]
                out["coll"][k, j] = int(cc[i])
                out["offroad"][k, j] = bool(a.is_offroad())
                ahead = a.obs.ahead_vehicle
                if ahead is not None:
                    out["lead_valid"][k, j] = 1
                    out["lead_dis"][k, j], out["lead_v"][k, j] = ahead[0], ahead[1]

    record(0)
    for k in range(n_ticks):
        eng.step()
        record(k + 1)
    ring = np.zeros((A, 11, 5))
    ring_valid = np.zeros(A, np.int32)
    for aid, a in am.agents.items():
        i = index_of[aid]
        hs = a.history_states
        ring[i, :, 0:2], ring[i, :, 2:4], ring[i, :, 4] = hs.xyz[:, :2], hs.vel, hs.heading[:, 0]
        ring_valid[i] = int(hs.valid.sum())
    out["ring"], out["ring_valid"] = ring, ring_valid
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, kw in CASES.items():
        kw = dict(kw)
        n_ticks = kw.pop("n_ticks")
        raw = city.make_city(**kw)
        with tempfile.TemporaryDirectory() as d:
            city.write_city(d, raw)
            ref = run_reference_city(d, n_ticks)
        st = city.pack_city(raw)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **{"ref_" + k: v for k, v in ref.items()},
                            **{"st_" + f.name: np.asarray(getattr(st, f.name)) for f in dataclasses.fields(st)})
        fin = int((ref["status"][-1] == 3).sum())
        print(name, "agents", len(raw["agents"]), "lanes", len(raw["lanes"]), "finished", fin, "max active", int(ref["n_active"].max()),
              "collisions", int(ref["coll"].sum()), "offroad", int(ref["offroad"].sum()), "lc", int(ref["is_lc"].sum()),
              "lead", int(ref["lead_valid"].sum()))


if __name__ == "__main__":
    main()
