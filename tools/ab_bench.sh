#!/bin/bash
# A/B builds on the GPU box: one bench line per set of -D defines (development only).  usage: tools/ab_bench.sh "" "-DX=1" "-DX=2"
mkdir -p gpurun_out
i=0
for defs in "$@"; do
  touch adaptive_classifier_b200/csrc/*.cu
  AC_NVCC_DEFS="$defs" python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
  timeout 900 python bench.py --no-extras --steps ${STEPS:-20} 2> /dev/null > gpurun_out/ab_$i.json
  python - "$defs" gpurun_out/ab_$i.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
r = d["roofline"]
print(f"[{sys.argv[1]}] {d['value']:.0f} q/s  {d['ms_per_step']:.3f} ms  gemm {r['achieved']:.0f} TF/s ({r['ms_total'] / d['steps']:.2f} ms/step)  "
      f"attn {d['attention']['us_per_layer']:.1f} us  knn {d['roofline_knn']['ms_per_launch']:.3f} ms  clk {d['clocks']['sm_mhz']}")
PY
  i=$((i+1))
done
touch adaptive_classifier_b200/csrc/*.cu
