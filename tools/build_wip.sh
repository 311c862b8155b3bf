#!/bin/sh
# Round-2 work in progress (csrc/wip/), kept OUT of librnnt_b200.so:
#   1. syntax / ptxas check of the 2-CTA kernel draft (joint_tc4.cuh) -- compiles, links nothing
#   2. csrc/wip/libwip.so with the 2-CTA MMA probe (export rnntb200_wip_mma2_probe) for tools/mma2_probe.py
set -e
cd "$(dirname "$0")/../rnnt_speech_recognition_b200/csrc"
unset CC CXX
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
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v -diag-suppress 177 -I. -I../../include -c /tmp/wip_check.cu -o /tmp/wip_check.o 2>&1 | grep -E "error|joint_tc4" | head -20
cuobjdump -sass /tmp/wip_check.o | grep -o "UTCHMMA[.A-Z0-9_]*\|UTCBAR[.A-Z0-9_]*\|UTMALDG[.A-Z0-9_]*\|UCGABAR[._A-Z0-9]*" | sort | uniq -c
cat > /tmp/wip_probe.cu <<'EOF'
#include <cuda_runtime.h>
#include "wip/mma2_probe.cuh"
extern "C" int rnntb200_wip_mma2_probe(int mode, int iters, int clusters, float* out_dev) {
    rb::mma2_probe_kernel<<<2 * clusters, 128, rb::c2::PROBE_KB * 4096 + 1024>>>(mode, iters, out_dev);
    return (int)cudaDeviceSynchronize();
}
EOF
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -I. -I../../include /tmp/wip_probe.cu -o wip/libwip.so
echo built wip/libwip.so
