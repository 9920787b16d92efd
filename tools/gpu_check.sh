#!/bin/bash
# One GPU call: the GPU suite in separate processes per area (a trapping kernel poisons only its own process), smoke, bench.
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider "$@" > gpurun_out/pytest_$name.log 2>&1; echo "$name: $(tail -1 gpurun_out/pytest_$name.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_$name.log | cut -c1-220 | head -20; }
run knn tests/test_gpu_parity.py -k "knn or golden_router or sharded or proto or segment"
run head tests/test_gpu_parity.py -k "head or ewc"
run encoder tests/test_gpu_parity.py -k "encoder or linear or pipeline"
run classifier tests/test_gpu_classifier.py tests/test_gpu_zz_threads.py
run training tests/test_gpu_training_golden.py
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "=== bench"; timeout 1200 python bench.py --cfg4-examples ${CFG4:-10240} 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-600; tail -5 gpurun_out/bench.err
