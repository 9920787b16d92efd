#!/bin/bash
# builds tools/pair_harness (torch-free C-ABI check of the opt-in CTA-pair kernels); the binary is git-ignored and
# travels to the GPU box with the gpurun snapshot, next to adaptive_classifier_b200/libadaptive_b200.so
set -e
cd "$(dirname "$0")"
python -c "import sys; sys.path.insert(0, '..'); from adaptive_classifier_b200 import build; print(build.build_library())"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lineinfo pair_harness.cu -o pair_harness \
    -L../adaptive_classifier_b200 -ladaptive_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../adaptive_classifier_b200'
echo built tools/pair_harness
