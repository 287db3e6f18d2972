"""N > 1 path on CPU: world_size-2 gloo.  Each rank steps its shard of the scenario list (host emulation of the
kernel — no GPU here) and the only collective of the path, the all-reduce of the int64 episode statistics, must
reproduce the single-process totals."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_scen, ticks, out_path):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    from hostemu import NumpyEmuBackend
    from lcsim_b200 import shard, workload
    from lcsim_b200.batch import BatchEngine

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    s0, s1 = shard.shard_range(n_scen, rank, world)
    sb = workload.synthetic_batch(400 + s0, s1 - s0, 48, workers=1)
    be = BatchEngine(sb, reactive=True, with_metrics=True, backend=NumpyEmuBackend())
    be.rollout(ticks)
    local = torch.from_numpy(np.asarray(be.episode_stats()).copy())
    total = shard.reduce_episode_stats(local)
    if rank == 0:
        np.save(out_path, np.stack([local.numpy(), total.numpy()]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_the_batch():
    from lcsim_b200 import shard

    for n in (1, 7, 8, 1024, 1025):
        for w in (1, 2, 3, 8):
            r = [shard.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(8, 2, 2)


def test_two_rank_gloo_stats_reduce(tmp_path):
    sys.path.insert(0, HERE)
    from hostemu import NumpyEmuBackend, build
    from lcsim_b200 import shard, workload
    from lcsim_b200.batch import BatchEngine

    build()
    n_scen, ticks = 5, 25
    out = str(tmp_path / "stats.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_scen, ticks, out), nprocs=2, join=True)
    got = np.load(out)
    sb = workload.synthetic_batch(400, n_scen, 48, workers=1)
    be = BatchEngine(sb, reactive=True, with_metrics=True, backend=NumpyEmuBackend())
    be.rollout(ticks)
    want = np.asarray(be.episode_stats())
    np.testing.assert_array_equal(got[1], want)
    assert got[0][0] < want[0]  # rank 0 alone saw only its shard
    r = shard.rates(torch.from_numpy(want))
    assert r["agent_steps"] == int(want[0]) and r["ticks"] == n_scen * ticks
