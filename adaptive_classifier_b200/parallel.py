"""Row-sharded prototype search across GPUs; one process per GPU, torch.distributed.

    E   data-parallel: every rank encodes its own B queries (replicated encoder weights, no communication)
    K   P is row-sharded (N/G contiguous rows per rank, global id = local + offset):
          1. all-gather of the unit embeddings      -> every rank holds all G*B queries   (G*B*D*4 bytes)
          2. local top-k over the shard for all G*B queries (ac_knn_l2_topk, row_offset = shard start)
          3. all-to-all of the per-shard candidates -> rank r receives the G lists of ITS B queries
          4. ac_topk_merge by (d, global id)        -> bit-identical to a single-shard search
    H   data-parallel on the rank's own queries; blend as in predict_batch.

The search / merge callables are injectable so the host logic (sharding arithmetic, collectives, merge order)
is covered by world_size-2 gloo tests on CPU with the oracle standing in for the kernels (tests only); the
default callables are the CUDA kernels and raise without a GPU.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(N: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous row range [lo, hi) of `rank`; the first N % world ranks get one extra row"""
    base, rem = divmod(N, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _cuda_search(Q, P, k, row_offset):
    from . import _cabi
    return _cabi.knn_l2_topk(Q, P, k, row_offset=row_offset)


def _cuda_merge(d, i):
    from . import _cabi
    return _cabi.topk_merge(d, i)


class ShardedIndex:
    """This rank's shard of the prototype matrix plus the collective search."""

    def __init__(self, P_local: torch.Tensor, row_offset: int, *, group=None,
                 search: Callable = _cuda_search, merge: Callable = _cuda_merge):
        self.P = P_local
        self.row_offset = int(row_offset)
        self.group = group
        self.search = search
        self.merge = merge
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def search_local_queries(self, q_local: torch.Tensor, k: int):
        """q_local [B, D] (this rank's queries) -> (d [B,k], global ids [B,k]) over the WHOLE index."""
        G = self.world
        B, D = q_local.shape
        if G == 1:
            return self.search(q_local, self.P, k, self.row_offset)
        q_all = torch.empty((G * B, D), dtype=q_local.dtype, device=q_local.device)
        dist.all_gather_into_tensor(q_all, q_local.contiguous(), group=self.group)
        d_loc, i_loc = self.search(q_all, self.P, k, self.row_offset)         # [G*B, k]
        d_recv = torch.empty_like(d_loc)
        i_recv = torch.empty_like(i_loc)
        # chunk g of the send buffer (queries of rank g) goes to rank g; received chunk g = shard g's list of MY queries
        dist.all_to_all_single(d_recv, d_loc.contiguous(), group=self.group)
        dist.all_to_all_single(i_recv, i_loc.contiguous(), group=self.group)
        return self.merge(d_recv.view(G, B, k), i_recv.view(G, B, k))
