#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu (single process)"; timeout 1200 python -m pytest tests/ -x -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
echo "=== bench"; timeout 900 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-300
