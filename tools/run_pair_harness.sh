#!/bin/bash
# GPU-side runner of tools/pair_harness (torch-free, C ABI only): every test in its own process, so a trapping kernel
# kills only its own log.  Round 1 ran the first five (profiles/r01_pair_*.log); defer / defer_full / epoch check the
# variants written after the round-1 GPU budget was spent.
mkdir -p gpurun_out
cd tools
for t in "linear 1" "knn 1" "encoder 1" "linear 2" "knn 2" "defer" "defer_full" "defer_layers" "epoch" "epi16" "attn" "pdl" "knn_epi" "cls_attn"; do
    name=${t// /_}
    timeout 40 ./pair_harness $t > ../gpurun_out/pair_${name}.log 2>&1
    echo "exit=$?" >> ../gpurun_out/pair_${name}.log
done
cd ..
tail -n 30 gpurun_out/pair_*.log
