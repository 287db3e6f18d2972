"""Synthetic WOMD-shaped batches for tests, smoke and bench: seed = scenario index."""
from __future__ import annotations

import os
from concurrent.futures import ProcessPoolExecutor

from . import packer, synth


def _one(args):
    seed, n_agents = args
    return packer.scenario_from_raw(synth.make_raw_scenario(seed, n_agents))


def synthetic_batch(seed0: int, n_scenarios: int, n_agents: int = 128, workers: int | None = None, **pack_kw):
    """StaticBatch of scenarios with seeds seed0 .. seed0+n-1 (generated in parallel when worthwhile)."""
    jobs = [(seed0 + i, n_agents) for i in range(n_scenarios)]
    if workers is None:
        workers = min(os.cpu_count() or 1, 32)
    if n_scenarios < 8 or workers <= 1:
        scs = [_one(j) for j in jobs]
    else:
        with ProcessPoolExecutor(max_workers=workers) as ex:
            scs = list(ex.map(_one, jobs, chunksize=max(1, n_scenarios // (4 * workers))))
    return packer.pack(scs, **pack_kw)
