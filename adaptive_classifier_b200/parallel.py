"""Row-sharded prototype search across GPUs; one process per GPU, torch.distributed.

    E   data-parallel: every rank encodes its own B queries (replicated encoder weights, no communication)
    K   P is row-sharded (N/G contiguous rows per rank, global id = local + offset):
          1. all-gather of the unit embeddings      -> every rank holds all G*B queries   (G*B*D*4 bytes)
          2. local top-k over the shard for all G*B queries (ac_knn_l2_topk, row_offset = shard start)
          3. all-to-all of the per-shard candidates -> rank r receives the G lists of ITS B queries
          4. ac_topk_merge by (d, global id)        -> bit-identical to a single-shard search
    H   data-parallel on the rank's own queries; blend as in predict_batch.

Exchange over peer memory (opt-in, `ShardedIndex(..., exchange=PeerExchange(...))`; csrc/peer.cu): steps 1 and 3 become
stores into NVLink-mapped buffers of the consumers (torch symmetric memory provides the mapping) plus sequence-number
flags, instead of three NCCL collectives: the encoder's last kernel (or one scatter kernel) writes the embeddings to all
peers, one scatter kernel writes every shard's candidate block to the rank that owns those queries.

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


class PeerExchange:
    """Double-buffered symmetric exchange buffers + flags of one rank, mapped on every peer (GPU only).

    Layout of the symmetric allocation (identical on all ranks), for parity par in {0, 1}:
        q_all[par]   [G, B, D] fp32    slot g <- unit embeddings of rank g            (written by rank g on every rank)
        cand_d[par]  [G, B, k] fp32    slot g <- shard g's distances for MY queries    (written by rank g)
        cand_i[par]  [G, B, k] int64   slot g <- shard g's global ids for MY queries
        flags        [3, G]    uint32  channel 0: embeddings, 1: distances, 2: ids; entry g = last sequence number rank g
                                       has completely stored here
    Why two buffers are enough: rank A can only overwrite buffer `par` of rank B at step t+2; A reaches its step t+2 stores
    only after its own merge of step t+1, which waited for B's candidates of step t+1, which B produced after (stream
    order) everything it read from buffer `par` in step t.  Flags carry the step number, so a fast peer's later store never
    satisfies an earlier wait.
    """

    def __init__(self, B: int, D: int, k: int, group=None, device=None, *, backend=None):
        """`backend` (tests only) replaces torch's symmetric memory, the process group and the C-ABI binding with stand-ins:
        an object with .world, .rank, .symm (empty / rendezvous), .cabi (peer_table / peer_scatter / peer_wait), .device."""
        if backend is None:
            import torch.distributed._symmetric_memory as symm
            from . import _cabi
            self.group = group if group is not None else dist.group.WORLD
            self.G, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
            dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        else:
            symm, _cabi, self.group, self.G, self.rank, dev = backend.symm, backend.cabi, group, backend.world, backend.rank, backend.device
        self._cabi = _cabi
        self.B, self.D, self.k = B, D, k
        G = self.G
        al = lambda n: (n + 255) // 256 * 256
        self.q_bytes = al(G * B * D * 4)
        self.d_bytes = al(G * B * k * 4)
        self.i_bytes = al(G * B * k * 8)
        self.par_bytes = self.q_bytes + self.d_bytes + self.i_bytes
        self.flags_off = 2 * self.par_bytes
        total = self.flags_off + al(3 * G * 4)
        self.buf = symm.empty(total, dtype=torch.uint8, device=dev)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, self.group)
        if backend is None:
            torch.cuda.synchronize(dev)
        self.hdl.barrier()                                     # every rank's flags are zero before anyone stores
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.base_ptr = ptrs[self.rank]                        # == self.buf.data_ptr() on a real device
        self.tables = [_cabi.peer_table(G, self.rank, ptrs, [p + self.flags_off + ch * G * 4 for p in ptrs]) for ch in range(3)]
        self.counter = torch.zeros((4,), dtype=torch.int32, device=dev)
        self.seq = 0

    # offsets inside the symmetric buffer
    def q_off(self, par: int) -> int:
        return par * self.par_bytes

    def d_off(self, par: int) -> int:
        return par * self.par_bytes + self.q_bytes

    def i_off(self, par: int) -> int:
        return par * self.par_bytes + self.q_bytes + self.d_bytes

    def _view(self, off: int, nbytes: int, dtype, shape):
        return self.buf[off : off + nbytes].view(dtype).view(*shape)

    def next_step(self):
        """-> (sequence number, parity) of the step about to run"""
        self.seq += 1
        return self.seq, self.seq & 1

    def gather_queries(self, q_local: torch.Tensor, seq: int, par: int, already_scattered: bool = False) -> torch.Tensor:
        """every rank's [B, D] embeddings -> [G*B, D] (rank-major) on this rank"""
        c, G, B, D = self._cabi, self.G, self.B, self.D
        if not already_scattered:
            c.peer_scatter(q_local.contiguous(), B * D * 4, False, self.tables[0], self.q_off(par) + self.rank * B * D * 4, seq, self.counter[0:1])
        c.peer_wait(self.base_ptr + self.flags_off, G, seq)
        return self._view(self.q_off(par), G * B * D * 4, torch.float32, (G * B, D))

    def exchange_candidates(self, d_loc: torch.Tensor, i_loc: torch.Tensor, seq: int, par: int):
        """d_loc / i_loc [G*B, k] (block g = queries of rank g over MY shard) -> ([G, B, k], [G, B, k]) for MY queries"""
        c, G, B, k = self._cabi, self.G, self.B, self.k
        c.peer_scatter(d_loc.contiguous(), B * k * 4, True, self.tables[1], self.d_off(par) + self.rank * B * k * 4, seq, self.counter[1:2])
        c.peer_scatter(i_loc.contiguous(), B * k * 8, True, self.tables[2], self.i_off(par) + self.rank * B * k * 8, seq, self.counter[2:3])
        c.peer_wait(self.base_ptr + self.flags_off + G * 4, 2 * G, seq)
        return (self._view(self.d_off(par), G * B * k * 4, torch.float32, (G, B, k)),
                self._view(self.i_off(par), G * B * k * 8, torch.int64, (G, B, k)))


class ShardedIndex:
    """This rank's shard of the prototype matrix plus the collective search."""

    def __init__(self, P_local: torch.Tensor, row_offset: int, *, group=None,
                 search: Callable = _cuda_search, merge: Callable = _cuda_merge, exchange: Optional["PeerExchange"] = None):
        self.exchange = exchange
        self.P = P_local
        self.row_offset = int(row_offset)
        self.group = group
        self.search = search
        self.merge = merge
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if exchange is not None:                               # the exchange knows its world (tests inject one without torch.distributed)
            self.world, self.rank = exchange.G, exchange.rank

    def search_local_queries(self, q_local: torch.Tensor, k: int):
        """q_local [B, D] (this rank's queries) -> (d [B,k], global ids [B,k]) over the WHOLE index."""
        G = self.world
        B, D = q_local.shape
        if G == 1:
            return self.search(q_local, self.P, k, self.row_offset)
        if self.exchange is not None:
            return self._search_peer(q_local, k)
        q_all = torch.empty((G * B, D), dtype=q_local.dtype, device=q_local.device)
        dist.all_gather_into_tensor(q_all, q_local.contiguous(), group=self.group)
        d_loc, i_loc = self.search(q_all, self.P, k, self.row_offset)         # [G*B, k]
        d_recv = torch.empty_like(d_loc)
        i_recv = torch.empty_like(i_loc)
        # chunk g of the send buffer (queries of rank g) goes to rank g; received chunk g = shard g's list of MY queries
        dist.all_to_all_single(d_recv, d_loc.contiguous(), group=self.group)
        dist.all_to_all_single(i_recv, i_loc.contiguous(), group=self.group)
        return self.merge(d_recv.view(G, B, k), i_recv.view(G, B, k))

    def _search_peer(self, q_local: torch.Tensor, k: int, step=None, already_scattered: bool = False):
        """the same four steps with the exchange done by stores into peer memory (PeerExchange)"""
        ex = self.exchange
        assert k == ex.k and q_local.shape == (ex.B, ex.D)
        seq, par = step if step is not None else ex.next_step()
        q_all = ex.gather_queries(q_local, seq, par, already_scattered)
        d_loc, i_loc = self.search(q_all, self.P, k, self.row_offset)         # [G*B, k]
        d_recv, i_recv = ex.exchange_candidates(d_loc, i_loc, seq, par)
        return self.merge(d_recv, i_recv)
