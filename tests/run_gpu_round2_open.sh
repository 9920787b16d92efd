#!/bin/bash
# FIRST GPU call of round 2 (about 20 minutes of box time; trim the bench variant list in step 3 if the budget is tight): everything written after the round-1 GPU budget ran out is
# checked and measured here in one go.  Nothing in this script changes defaults; it only produces evidence in gpurun_out/.
#   1. torch-free C-ABI harness: pair kernels (re-check), deferred LayerNorm, fused head epoch, 16 epilogue warps, pipelined
#      attention, per-lane kNN epilogue, programmatic dependent launch
#   2. the GPU suite with the experimental tests enabled
#   3. bench.py: default, and with each variant / the combination switched on through AC_OPTIONS
#   4. config 4 (add_examples loop) with and without the fused epoch
#   5. ncu: launch list of the best combination + one --set full capture of the encoder GEMMs under ln_defer
mkdir -p gpurun_out
echo "=== 1. harness"; bash tools/run_pair_harness.sh > gpurun_out/harness.log 2>&1; grep -h "VARIANT\|PAIR ==\|FUSED ==\|MISMATCH\|TOLERANCE\|WITHIN\|watchdog\|error\|exit=" gpurun_out/pair_*.log | sort | uniq -c
echo "=== 2. pytest -m gpu with experimental variants"
AC_TEST_EXPERIMENTAL=1 timeout 1500 python -m pytest tests/ -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_experimental.log
echo "=== 3. bench variants"
for v in "" "knn_epi=1" "cls_attn=1" "ln_defer=1" "epi16=1" "epi16=3" "attn_pipe=1" "ln_defer=1,epi16=1,attn_pipe=1" "gemm_pair=1,ln_defer=1,epi16=1,attn_pipe=1" "gemm_pair=1,ln_defer=1,epi16=1,attn_pipe=1,pdl=1,knn_epi=1,cls_attn=1"; do
    name=$(echo "${v:-default}" | tr ',=' '__')
    AC_OPTIONS="$v" timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/bench_${name}.err | tee gpurun_out/bench_${name}.json | python -c "
import json,sys
j=json.loads(sys.stdin.readline()); print('$name', round(j['value']), 'q/s', round(j['ms_per_step'],3), 'ms/step; gemm', round(j['roofline']['achieved']), 'TFLOP/s share', round(j['roofline']['share_of_step'],3), '; e2e', round(j['e2e']['value']))"
done
echo "=== 4. add_examples loop"
for v in "" "head_fused=1"; do
    name=$(echo "${v:-default}" | tr ',=' '__')
    AC_OPTIONS="$v" timeout 900 python tools/bench_add_examples.py --examples 5120 2>&1 | tail -1 | tee gpurun_out/bench_add_examples_${name}.json | cut -c1-420
done
echo "=== 5. ncu"
AC_OPTIONS="ln_defer=1,epi16=1,attn_pipe=1" timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02_defer.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
AC_OPTIONS="ln_defer=1,epi16=1" timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 40 -c 8 -o gpurun_out/r02_gemm_defer \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -30
