#!/bin/bash
# GPU-side runner of tools/pair_harness: every test in its own process (a trapping kernel kills only its own log)
mkdir -p gpurun_out
cd tools
for t in "linear 1" "knn 1" "encoder 1" "linear 2" "knn 2"; do
    name=${t// /_}
    timeout 25 ./pair_harness $t > ../gpurun_out/pair_${name}.log 2>&1
    echo "exit=$?" >> ../gpurun_out/pair_${name}.log
done
cd ..
tail -n 30 gpurun_out/pair_*.log
