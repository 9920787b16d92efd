#!/bin/bash
# multi-GPU sanity: bench.py under torchrun on N GPUs (gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
echo "=== diag head"; timeout 300 python tests/diag_head.py 2>&1 | tail -20
echo "=== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | cut -c1-1200
tail -5 gpurun_out/bench_n$N.err | cut -c1-300
echo "=== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-400
echo "=== peer-memory exchange vs NCCL (csrc/peer.cu)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    tests/peer_exchange_check.py 2>&1 | tail -5 | tee gpurun_out/peer_exchange_n$N.log
echo "=== bench N=$N with AC_EXCHANGE=peer"
AC_EXCHANGE=peer timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/bench_peer_n$N.err | tee gpurun_out/bench_peer_n$N.json | cut -c1-400
tail -3 gpurun_out/bench_peer_n$N.err | cut -c1-300
