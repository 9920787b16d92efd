#!/bin/bash
# smoke + bench + ncu captures (run under gpurun; results land in gpurun_out/)
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-1500
tail -5 gpurun_out/bench.err
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee gpurun_out/bench_ref.json | cut -c1-600
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rows 200000 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
echo "=== ncu full: encoder GEMM"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 4 -f -o gpurun_out/prof_gemm \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rows 200000 > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log | cut -c1-300
echo "=== kNN only"; timeout 300 python tests/prof_knn.py 2>&1 | tail -2
echo "=== config 4: add_examples loop"; timeout 600 python tools/bench_add_examples.py --examples 5120 2>&1 | tail -1 | tee gpurun_out/bench_add_examples.json | cut -c1-700
echo "=== ncu full: kNN scan"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 2 -f -o gpurun_out/prof_knn \
    python tests/prof_knn.py > gpurun_out/ncu_knn_only.log 2>&1
tail -2 gpurun_out/ncu_knn_only.log | cut -c1-300
echo "=== ncu full: kNN coarse + attention"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attention_kernel|layernorm_kernel" -s 30 -c 2 -f -o gpurun_out/prof_knn_att \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_knn.log 2>&1
tail -2 gpurun_out/ncu_knn.log | cut -c1-300
ls -la gpurun_out
