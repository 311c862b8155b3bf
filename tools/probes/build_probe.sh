#!/bin/sh
# Builds tools/probes/libprobe.so with the 2-CTA tcgen05.mma bring-up probe (export rnntb200_wip_mma2_probe) for
# tools/probes/mma2_probe.py.  Outside the library build and outside the package tree.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
CSRC="$HERE/../../rnnt_speech_recognition_b200/csrc"
unset CC CXX
cat > /tmp/probe_main.cu <<'EOF2'
#include <cuda_runtime.h>
#include "mma2_probe.cuh"
extern "C" int rnntb200_wip_mma2_probe(int mode, int iters, int clusters, float* out_dev) {
    rb::mma2_probe_kernel<<<2 * clusters, 128, rb::c2::PROBE_KB * 4096 + 1024>>>(mode, iters, out_dev);
    return (int)cudaDeviceSynchronize();
}
EOF2
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -I"$HERE" -I"$CSRC" -I"$CSRC/../../include" /tmp/probe_main.cu -o "$HERE/libprobe.so"
echo built "$HERE/libprobe.so"
