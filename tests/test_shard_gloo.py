"""CPU test of the N>1 host logic: 2 ranks over gloo shard a query list and gather hit records."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_queries, out_path):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from usearch12_b200 import shard
    from usearch12_b200.capi import HIT_DTYPE
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(n_queries, rank, world)
    # fake local results: every third query of the shard has a hit on target (global query % 7)
    q = np.arange(lo, hi)
    sel = q[q % 3 == 0]
    hits = np.zeros(len(sel), dtype=HIT_DTYPE)
    hits["query"] = sel - lo
    hits["target"] = sel % 7
    hits["ids"] = 200 + rank
    merged = shard.gather_hits(hits, lo, rank, world)
    if rank == 0:
        np.save(out_path, merged)
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    from usearch12_b200 import shard
    for n in (0, 1, 7, 100, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_gather_hits_two_ranks_gloo(tmp_path):
    n = 1001
    out = str(tmp_path / "merged.npy")
    mp.spawn(_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
    merged = np.load(out)
    want = np.arange(n)[np.arange(n) % 3 == 0]
    assert np.array_equal(merged["query"], want)
    assert np.array_equal(merged["target"], want % 7)
    assert set(merged["ids"]) == {200, 201}
