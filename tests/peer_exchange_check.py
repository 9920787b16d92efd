"""Multi-GPU check of the peer-memory exchange (csrc/peer.cu, parallel.PeerExchange) against the NCCL exchange.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/peer_exchange_check.py

Every rank searches its own queries over a row-sharded index for several steps through both exchanges and demands
bit-identical (d, global id); ranks are deliberately skewed in time (sleeps) so that a fast rank runs a step ahead and both
buffer parities and the sequence-number flags are exercised.  Not a pytest test: it needs >= 2 GPUs (gpurun --gpus 2).
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from adaptive_classifier_b200 import _cabi
    from adaptive_classifier_b200.parallel import PeerExchange, ShardedIndex, shard_bounds
    N, D, B, k, steps = 200_000, 768, 64, 5, 12
    g = torch.Generator().manual_seed(0)
    P_all = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=1)
    lo, hi = shard_bounds(N, rank, world)
    P = P_all[lo:hi].to(dev).contiguous()
    search = lambda Q_, P_, k_, off_: _cabi.knn_l2_topk(Q_, P_, k_, row_offset=off_)
    nccl = ShardedIndex(P, lo, search=search)
    peer = ShardedIndex(P, lo, search=search, exchange=PeerExchange(B, D, k, device=dev))
    ok = True
    for step in range(steps):
        gq = torch.Generator().manual_seed(1000 * step + rank)
        Q = torch.nn.functional.normalize(P_all[torch.randint(0, N, (B,), generator=gq)] + 0.05 * torch.randn(B, D, generator=gq), dim=1).to(dev)
        if (step + rank) % 3 == 0:
            time.sleep(0.05)                      # skew the ranks: somebody is always a step ahead
        d_p, i_p = peer.search_local_queries(Q, k)
        d_p, i_p = d_p.clone(), i_p.clone()
        d_n, i_n = nccl.search_local_queries(Q, k)
        torch.cuda.synchronize()
        same = bool(torch.equal(d_p, d_n) and torch.equal(i_p, i_n))
        ok &= same
        if not same:
            print(f"rank {rank} step {step}: MISMATCH ({int((i_p != i_n).sum())} ids differ)", flush=True)
    # timing of the two exchanges around the same search (device time, max over ranks)
    def timed(index):
        Q = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1)
        for _ in range(3):
            index.search_local_queries(Q, k)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            index.search_local_queries(Q, k)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    t_n, t_p = timed(nccl), timed(peer)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"peer exchange vs NCCL over {steps} skewed steps on {world} GPUs: {'BIT-IDENTICAL' if int(flag) else 'MISMATCH'}; "
              f"sharded search per call: NCCL {t_n:.3f} ms, peer memory {t_p:.3f} ms", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
