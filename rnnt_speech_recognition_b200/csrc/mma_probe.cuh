// mma_probe.cuh -- bring-up microbenchmark: sustained cycles per tcgen05.mma for the shapes / operand sources
// the joint kernels use.  One elected lane of warp 0 issues `iters` MMAs back to back (operands are zeros in
// smem / TMEM), commits once, and the CTA reports (clock64 delta) / iters.  Exported for tools/mma_probe.py.
#pragma once
#include "ptx.cuh"

namespace rb {

__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int variant, int iters, float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (warp == 0) { ptx::tmem_alloc(&tmem_ptr, 512); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    const int N = (variant == 0 || variant == 5) ? 256 : ((variant == 1 || variant == 4) ? 128 : 64);
    const bool ts = variant >= 3;
    const bool alt = variant == 6;
    const bool walk = variant >= 7;   // 9: + a tcgen05.commit every 40 MMAs; 10: + other warps stream tcgen05.ld of a different accumulator
    __shared__ uint64_t bars2[8];
    if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) ptx::mbar_init(&bars2[i], 1);
    __syncthreads();   // operands walk like in the joint kernel: A over 320 TMEM columns, B over 40 KB
    const uint32_t idesc = ptx::umma_idesc_bf16(128, N);
    const uint64_t ad = ptx::umma_desc_k_sw128(ptx::smem_u32(smem));
    const uint64_t bd = ptx::umma_desc_k_sw128(ptx::smem_u32(smem + 32768));
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        t0 = clock64();
        if (walk) {
            for (int i = 0; i < iters; i += 20) {
                if (ptx::elect_one()) {
#pragma unroll
                    for (int j = 0; j < 5; ++j)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t a = tmem + ((variant == 8) ? 0 : (((i / 20) & 1) * 160)) + j * 32 + k * 8;
                            ptx::umma_bf16_ts(tmem + 320 + (((i / 40) % 3) * 64), a, bd + (uint64_t)(j * 512 + k * 2), idesc, 1u);
                        }
                    if (variant >= 9 && ((i / 20) & 1)) ptx::umma_commit(&bars2[(i / 40) & 7]);
                }
                __syncwarp();
            }
        } else
        for (int i = 0; i < iters; i += 4) {
            if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t d = tmem + 256 + ((alt && (k & 1)) ? 64 : 0);
                    if (ts) ptx::umma_bf16_ts(d, tmem + k * 8, bd + (uint64_t)(k * 2), idesc, 1u);
                    else ptx::umma_bf16(d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, 1u);
                }
            }
            __syncwarp();
        }
        if (ptx::elect_one()) ptx::umma_commit(&bar);
        __syncwarp();
        ptx::mbar_wait(&bar, 0);
        t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = (float)(t1 - t0) / (float)iters;
    }
    if (variant == 10 && warp > 0) {
        float acc = 0.f;
        for (int it = 0; it < iters / 40; ++it) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + 448, v);
            ptx::tmem_ld_wait();
            acc += __uint_as_float(v[0]);
            ptx::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + 480, v);
            ptx::tmem_ld_wait();
            acc += __uint_as_float(v[1]);
        }
        if (acc == 123.456f) out[0] = acc;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

}  // namespace rb
