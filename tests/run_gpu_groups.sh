#!/bin/bash
# Runs the GPU parity groups in separate processes so a trapped kernel cannot poison later groups.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
{
for grp in "linear_tc or linear_f16" "knn_exact" "tensor" "proto or sharded or segment" "head or ewc" "encoder" "pipeline or reduces"; do
  echo "=== group: $grp"
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -k "$grp" 2>&1 | tail -40
done
echo "=== file: test_gpu_classifier.py"
timeout 900 python -m pytest tests/test_gpu_classifier.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -60
} 2>&1 | tee gpurun_out/gpu_groups.log | tail -200
