#!/bin/sh
# Syntax / ptxas check of the round-2 work in progress (csrc/wip/): compiles, links nothing, ships nothing.
set -e
cd "$(dirname "$0")/../rnnt_speech_recognition_b200/csrc"
cat > /tmp/wip_check.cu <<'EOF'
#include <cuda_runtime.h>
#include "kernels_simt.cuh"
#include "timing.cuh"
#include "joint_tc.cuh"
#include "wip/joint_tc4.cuh"
rnntStatus_t wip_instantiate(const rb::Tc2Geom& g, const CUtensorMap& a, const rb::JointTcParams& p, cudaStream_t s) {
    rnntStatus_t st = rb::tc4_launch<0>(g, a, a, a, p, s);
    if (!st) st = rb::tc4_launch<1>(g, a, a, a, p, s);
    if (!st) st = rb::tc4_launch<2>(g, a, a, a, p, s);
    return st;
}
EOF
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v -diag-suppress 177 -I. -I../../include -c /tmp/wip_check.cu -o /tmp/wip_check.o 2>&1 | grep -E "error|joint_tc4|Used" | head -20
cuobjdump -sass /tmp/wip_check.o | grep -o "UTCHMMA[.A-Z0-9_]*\|UTCBAR[.A-Z0-9_]*\|UTMALDG[.A-Z0-9_]*\|UCGABAR[._A-Z0-9]*" | sort | uniq -c
