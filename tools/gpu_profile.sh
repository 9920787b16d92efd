#!/bin/bash
# ncu evidence for profiles/ (one GPU, never a multi-rank command): launch list of two steps, full captures of the dominant kernels
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2_kernel|attention_kernel|ln_stats_kernel|layernorm_kernel" -s 300 -c 9 -o gpurun_out/r02_encoder_full $B > gpurun_out/ncu_enc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|topk_chunk_kernel|knn_rerank_kernel|embed_ln_kernel|sgemm_nt_kernel" -s 20 -c 8 -o gpurun_out/r02_knn_full $B > gpurun_out/ncu_knn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"head_train_kernel" -s 1 -c 1 -o gpurun_out/r02_head_full python tools/head_phase_times.py 20 > gpurun_out/ncu_head.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches.csv
for f in enc knn head; do tail -n 2 gpurun_out/ncu_$f.log; done
