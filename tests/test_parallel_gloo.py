"""world_size-2 gloo test of the row-sharded search: sharding arithmetic, all-gather of the
queries, all-to-all of the per-shard candidates and the (d, id) merge give exactly the single-index result.
The oracle stands in for the CUDA kernels through ShardedIndex's injectable callables (tests only)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from adaptive_classifier_b200.parallel import ShardedIndex, shard_bounds
    from oracle import knn_oracle as ko
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, D, B, k = 3001, 32, 5, 4
    rng = np.random.default_rng(0)
    P = rng.standard_normal((N, D)).astype(np.float32)
    P[2000] = P[10]                                           # a cross-shard exact tie
    Qall = rng.standard_normal((world * B, D)).astype(np.float32)
    Qall[0] = P[10]
    lo, hi = shard_bounds(N, rank, world)

    def search(Q, Pl, kk, off):
        d, i = ko.knn_l2(Q.numpy(), Pl.numpy(), kk, row_offset=off)
        return torch.from_numpy(d), torch.from_numpy(i)

    def merge(d, i):
        od, oi = ko.topk_merge(d.numpy(), i.numpy())
        return torch.from_numpy(od), torch.from_numpy(oi)

    idx = ShardedIndex(torch.from_numpy(P[lo:hi]), lo, search=search, merge=merge)
    d, i = idx.search_local_queries(torch.from_numpy(Qall[rank * B : (rank + 1) * B]), k)
    d0, i0 = ko.knn_l2(Qall[rank * B : (rank + 1) * B], P, k)
    ok = np.array_equal(i.numpy(), i0) and np.array_equal(d.numpy(), d0)
    if rank == 0:
        ok = ok and i0[0, 0] == 10 and i0[0, 1] == 2000      # tie -> lower global id first
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_search_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r, False) for r in range(world)), dict(ret)
