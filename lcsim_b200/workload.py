"""Synthetic WOMD-shaped batches for tests, smoke and bench: seed = scenario index."""
from __future__ import annotations

import os
from concurrent.futures import ProcessPoolExecutor

from . import packer, synth


def _one(args):
    seed, n_agents, with_centres = args
    sc = packer.scenario_from_raw(synth.make_raw_scenario(seed, n_agents))
    if with_centres:  # the (P,3) polyline centres of the observation builder, computed where the lanes already are
        centres = packer.polyline_centres(sc.lanes)
        lines = packer.centreline_points(sc.lanes, sc.lines)[:2] if with_centres == "lanes" else None
        sc.lanes, sc.lines, sc.edges_raw = [], [], []
        return sc, centres, lines
    return sc


def synthetic_batch(seed0: int, n_scenarios: int, n_agents: int = 128, workers: int | None = None,
                    with_centres: bool = False, **pack_kw):
    """StaticBatch of scenarios with seeds seed0 .. seed0+n-1 (generated in parallel when worthwhile).
    with_centres: returns (batch, [polyline centres per scenario]) for BatchEngine.upload_polylines; with_centres="lanes"
    returns (batch, centres, [(xy_theta, ids) per scenario]) with the centreline points for upload_centrelines too."""
    jobs = [(seed0 + i, n_agents, with_centres) for i in range(n_scenarios)]
    if workers is None:
        workers = min(os.cpu_count() or 1, 32)
    if n_scenarios < 8 or workers <= 1:
        scs = [_one(j) for j in jobs]
    else:
        with ProcessPoolExecutor(max_workers=workers) as ex:
            scs = list(ex.map(_one, jobs, chunksize=max(1, n_scenarios // (4 * workers))))
    if with_centres == "lanes":
        return packer.pack([x[0] for x in scs], **pack_kw), [x[1] for x in scs], [x[2] for x in scs]
    if with_centres:
        return packer.pack([x[0] for x in scs], **pack_kw), [x[1] for x in scs]
    return packer.pack(scs, **pack_kw)
