#!/bin/bash
# what the driver runs at round end, in one go: the whole GPU suite in ONE process, smoke, bench
mkdir -p gpurun_out
echo "=== pytest -m gpu (single process)"; timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "=== config 4"; timeout 900 python tools/bench_add_examples.py --examples 5120 2>&1 | tail -1 | tee gpurun_out/bench_add_examples.json | cut -c1-500
echo "=== latency"; timeout 600 python tools/bench_latency.py 2>&1 | tail -1 | tee gpurun_out/bench_latency.json
echo "=== bench"; timeout 900 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-330
echo "=== cfg2"; timeout 600 python bench.py --workload cfg2 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_cfg2.json | cut -c1-330
echo "=== cfg5 (roberta-large shapes, 1 GPU)"; timeout 900 python bench.py --workload cfg5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_cfg5.json | cut -c1-330
