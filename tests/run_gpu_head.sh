#!/bin/bash
mkdir -p gpurun_out
for grp in "head or ewc" "pipeline"; do
  echo "=== group: $grp"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -k "$grp" 2>&1 | tail -6
done
echo "=== classifier"; timeout 600 python -m pytest tests/test_gpu_classifier.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -4
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "=== config 4"; timeout 900 python tools/bench_add_examples.py --examples 5120 2>&1 | tail -1 | tee gpurun_out/bench_add_examples.json | cut -c1-600
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-330
