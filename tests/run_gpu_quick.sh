#!/bin/bash
# quick iteration: kNN-related tests, smoke, kNN-only timing (+ ncu), bench
mkdir -p gpurun_out
for grp in "tensor" "pipeline or sharded" "head or ewc"; do
  echo "=== group: $grp"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -k "$grp" 2>&1 | tail -8
done
echo "=== classifier"; timeout 600 python -m pytest tests/test_gpu_classifier.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -8
echo "=== config 4"; timeout 600 python tools/bench_add_examples.py --examples 5120 2>&1 | tail -1 | tee gpurun_out/bench_add_examples.json | cut -c1-600
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "=== kNN only"; timeout 300 python tests/prof_knn.py 2>&1 | tail -2
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-400
tail -3 gpurun_out/bench.err
echo "=== ncu full: kNN scan"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 2 -f -o gpurun_out/prof_knn \
    python tests/prof_knn.py > gpurun_out/ncu_knn_only.log 2>&1
tail -2 gpurun_out/ncu_knn_only.log | cut -c1-300
