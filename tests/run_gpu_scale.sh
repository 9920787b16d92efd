#!/bin/bash
# scaling evidence on one 8-GPU box: N = 8, 4 (and 1 for the same-box ratio)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for N in 8 4; do
  echo "=== bench N=$N"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
      bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | cut -c1-260
  tail -2 gpurun_out/bench_n$N.err | cut -c1-200
done
echo "=== bench N=1"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_n1.json | cut -c1-260
