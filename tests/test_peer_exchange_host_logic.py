"""Host logic of parallel.PeerExchange / ShardedIndex._search_peer without GPUs: G threads stand in for the ranks, a fake
symmetric-memory module hands every "rank" the same list of CPU buffers, and the two C-ABI entries (ac_peer_scatter,
ac_peer_wait) are replaced by Python functions that copy bytes / spin on the flags exactly as csrc/peer.cu is specified to.
What is under test is the product's offset arithmetic, slot layout, double buffering and sequence numbers: every rank's
result must equal the single-index search, over several steps with deliberately skewed ranks.  (The kernels themselves are
covered by tests/test_cpu_emulated_kernels.py, the protocol by tests/test_sim_peer_exchange.py.)"""
import ctypes
import threading
import time

import numpy as np
import torch

from oracle import knn_oracle as ko


class FakeWorld:
    """shared state of the G fake ranks: one byte buffer per rank, addressed by fake 'device pointers'"""

    def __init__(self, G):
        self.G = G
        self.bufs = [None] * G
        self.base = [(r + 1) << 40 for r in range(G)]          # fake pointer of rank r's buffer
        self.barrier = threading.Barrier(G)
        self.lock = threading.Lock()

    def resolve(self, ptr):
        r = (ptr >> 40) - 1
        return self.bufs[r], ptr - self.base[r]


def _make_rank(world, rank, B, D, k, N, P_all, Qs, steps, results, errors):
    import adaptive_classifier_b200.parallel as par
    from adaptive_classifier_b200.parallel import PeerExchange, ShardedIndex, shard_bounds

    class Handle:
        def __init__(self, buf):
            world.bufs[rank] = buf
            self.buffer_ptrs = world.base

        def barrier(self, *a, **kw):
            world.barrier.wait()

    class FakeSymm:
        @staticmethod
        def empty(n, dtype=None, device=None):
            return torch.zeros(n, dtype=torch.uint8)

        @staticmethod
        def rendezvous(buf, group):
            h = Handle(buf)
            world.barrier.wait()                               # everybody has registered its buffer
            return h

    class FakeCabi:
        AC_MAX_PEERS = 16

        @staticmethod
        def peer_table(G, r, buf_ptrs, flag_ptrs):
            return {"world": G, "rank": r, "buf": list(buf_ptrs), "flag": list(flag_ptrs)}

        @staticmethod
        def peer_scatter(src, bytes_per_dst, blocks_mode, table, dst_offset, seq, counter):
            raw = src.contiguous().view(torch.uint8).reshape(-1)
            assert bytes_per_dst % 16 == 0 and dst_offset % 16 == 0
            for p in range(table["world"]):
                buf, off = world.resolve(table["buf"][p] + dst_offset)
                s0 = p * bytes_per_dst if blocks_mode else 0
                buf[off : off + bytes_per_dst] = raw[s0 : s0 + bytes_per_dst]
            for p in range(table["world"]):                    # data first, then the flags (release)
                buf, off = world.resolve(table["flag"][p] + 4 * table["rank"])
                buf[off : off + 4] = torch.tensor([seq & 0xFFFFFFFF], dtype=torch.int64).view(torch.uint8)[:4]

        @staticmethod
        def peer_wait(flags_ptr, n, seq):
            buf, off = world.resolve(flags_ptr)
            t0 = time.time()
            while True:
                flags = buf[off : off + 4 * n].clone().view(torch.int32)
                if bool((flags >= seq).all()):
                    return
                assert time.time() - t0 < 60, f"rank {rank} waits for seq {seq}, has {flags.tolist()}"
                time.sleep(0.0005)

    try:
        # the product's PeerExchange.__init__ with the fakes injected (no torch.distributed, no CUDA)
        class Backend:
            pass
        be = Backend()
        be.world, be.rank, be.symm, be.cabi, be.device = world.G, rank, FakeSymm, FakeCabi, torch.device("cpu")
        ex = PeerExchange(B, D, k, backend=be)

        lo, hi = shard_bounds(N, rank, world.G)

        def search(Q, Pl, kk, off):
            d, i = ko.knn_l2(Q.numpy().copy(), Pl.numpy(), kk, row_offset=off)
            return torch.from_numpy(d), torch.from_numpy(i)

        def merge(d, i):
            od, oi = ko.topk_merge(d.numpy().copy(), i.numpy().copy())
            return torch.from_numpy(od), torch.from_numpy(oi)

        idx = ShardedIndex(torch.from_numpy(P_all[lo:hi]), lo, search=search, merge=merge, exchange=ex)
        assert (idx.world, idx.rank) == (world.G, rank)
        out = []
        for step in range(steps):
            if (step + rank) % 3 == 0:
                time.sleep(0.01)                               # skew: somebody is always ahead
            d, i = idx._search_peer(torch.from_numpy(Qs[step][rank]), k)
            out.append((d.numpy().copy(), i.numpy().copy()))
        results[rank] = out
    except Exception as e:                                      # surfaced by the main thread
        errors.append((rank, repr(e)))
        try:
            world.barrier.abort()
        except Exception:
            pass


def test_peer_exchange_offsets_and_double_buffering_with_fake_ranks():
    G, B, D, k, N, steps = 3, 8, 32, 4, 2001, 7
    rng = np.random.default_rng(0)
    P_all = rng.standard_normal((N, D)).astype(np.float32)
    P_all[1500] = P_all[7]                                      # a cross-shard exact tie
    Qs = [[rng.standard_normal((B, D)).astype(np.float32) for _ in range(G)] for _ in range(steps)]
    Qs[0][0][0] = P_all[7]
    world = FakeWorld(G)
    results, errors = {}, []
    threads = [threading.Thread(target=_make_rank, args=(world, r, B, D, k, N, P_all, Qs, steps, results, errors)) for r in range(G)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    assert sorted(results) == list(range(G))
    for r in range(G):
        for step in range(steps):
            d0, i0 = ko.knn_l2(Qs[step][r], P_all, k)
            d, i = results[r][step]
            assert np.array_equal(i, i0) and np.array_equal(d, d0), (r, step)
    assert results[0][0][1][0, 0] == 7 and results[0][0][1][0, 1] == 1500       # tie -> lower global id first
