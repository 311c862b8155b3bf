// mma2_probe.cuh -- bring-up probe for the 2-CTA form of the fused kernel's mainloop ("Round-2 plan"):
//   tcgen05.mma.cta_group::2.kind::f16, M = 256 (128 TMEM lanes in each CTA of the pair), N = 64, K = 16 per
//   instruction, A read from EACH CTA's own tensor memory, B split across the pair (vocabulary rows [32*rank, +32) of
//   the chunk sit in CTA `rank`'s shared memory at the same offset), D written to the same TMEM columns of both CTAs.
// It answers, on hardware, the questions the 2-CTA kernel depends on before that kernel is written:
//   mode 0  correctness of the operand/accumulator mapping: A and B are small integers generated in the kernel,
//           every CTA dumps its 128 x 64 accumulator; tools/mma2_probe.py compares with the exact integer GEMM
//   mode 1  cycles per cta_group::2 MMA when one elected lane of the LEADER issues 20 per block (the fused kernel's
//           issue pattern), A walking over 320 columns, B over 5 K-block slabs
// Status: run on B200 in round 2 (profiles/r02/mma2_probe.log): every accumulator entry exact for 1 and 74 pairs,
// 32.1 cycles per cta_group::2 MMA (the dispatch floor).  Outside the library build: tools/probes/build_probe.sh
// compiles it into tools/probes/libprobe.so (export rnntb200_wip_mma2_probe) for tools/probes/mma2_probe.py.
#pragma once
#include "ptx.cuh"

namespace rb {
namespace c2 {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMEM management for a CTA pair: the SAME warp index of BOTH CTAs executes these (cute/arch/tmem_allocator_sm100.hpp:116-181)
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_result)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[tmem of each CTA] . B[smem halves of both CTAs]^T ; issued by ONE thread of the leader CTA
// (cute/arch/mma_sm100_umma.hpp:634-672, without the optional disable-output-lane vector)
__device__ __forceinline__ void umma_bf16_ts2(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued cta_group::2 MMAs of this thread complete -> one arrive on the barrier at this shared-memory
// offset in every CTA of `cta_mask` (cutlass/arch/barrier.h:846-863)
__device__ __forceinline__ void umma_commit2_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(ptx::smem_u32(bar)), "h"(cta_mask) : "memory");
}

constexpr int PROBE_KB = 5;            // K blocks of 64 (H = 320 here; the fused kernel has 10)
constexpr int PROBE_N = 64;            // vocabulary rows per MMA (32 per CTA)

// exact small-integer operands: every product and every partial sum is exactly representable in bf16 / fp32
__host__ __device__ inline float probe_a(int cta, int r, int k) { return (float)(((r * 3 + k * 5 + cta * 7) % 7) - 3); }
__host__ __device__ inline float probe_b(int n, int k) { return (float)(((n * 2 + k) % 5) - 2); }

}  // namespace c2

// grid = 2 * clusters CTAs, cluster (2,1,1), 128 threads.  out: mode 0 -> [cluster][2][128][64] floats; mode 1 -> [cluster] floats
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma2_probe_kernel(int mode, int iters, float* out) {
    using namespace c2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // B half of this CTA: PROBE_KB slabs of [32 rows x 64 k] bf16, K-major SWIZZLE_128B (4 KB each, 1024-aligned)
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_ptr;
    const uint32_t rank = cluster_ctarank(), cluster = blockIdx.x >> 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = threadIdx.x;   // r = TMEM lane = row of this CTA's A / D
    if (threadIdx.x == 0) { ptx::mbar_init(&done_bar, 1); ptx::fence_barrier_init(); }
    if (warp == 0) { tmem_alloc2(&tmem_ptr, 512); tmem_relinquish2(); }
    // ---- B half: row n_local (0..31) = vocabulary row 32*rank + n_local; chunk c of row n at (c ^ (n & 7)) * 16
    for (int i = threadIdx.x; i < PROBE_KB * 32 * 8; i += 128) {
        const int kb = i / 256, n = (i >> 3) & 31, c = i & 7;
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = kb * 64 + c * 8 + j * 2;
            const float b0 = mode == 0 ? probe_b(32 * rank + n, k) : 0.f, b1 = mode == 0 ? probe_b(32 * rank + n, k + 1) : 0.f;
            w[j] = ptx::pack_bf16x2(b0, b1);
        }
        *reinterpret_cast<uint4*>(smem + kb * 4096 + n * 128 + ((c ^ (n & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    // ---- A: this CTA's 128 rows x (PROBE_KB * 64) k as bf16 pairs in TMEM columns [0, PROBE_KB * 32)
    for (int kb = 0; kb < PROBE_KB; ++kb)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            uint32_t zr[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int k = kb * 64 + hh * 32 + j * 2;
                zr[j] = mode == 0 ? ptx::pack_bf16x2(probe_a(rank, r, k), probe_a(rank, r, k + 1)) : 0u;
            }
            ptx::tmem_st_32x16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kb * 32 + hh * 16), zr);
        }
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    cluster_sync();                       // both CTAs' operands are in place
    ptx::tc_fence_after();
    const uint32_t acc = tmem + 320;      // accumulator columns (the fused kernel's first buffer)
    const uint32_t idesc = ptx::umma_idesc_bf16(256, PROBE_N);
    const uint64_t bd = ptx::umma_desc_k_sw128(ptx::smem_u32(smem));
    long long t0 = 0;
    if (rank == 0 && warp == 1) {         // leader CTA: one elected lane issues for the pair
        t0 = clock64();
        const int blocks = mode == 0 ? 1 : iters / 20;
        for (int b = 0; b < blocks; ++b) {
            if (ptx::elect_one()) {
#pragma unroll
                for (int j = 0; j < PROBE_KB; ++j)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ts2(acc, tmem + j * 32 + k * 8, bd + (uint64_t)(j * 256 + k * 2), idesc,
                                      mode == 0 ? (uint32_t)((j | k) != 0) : 1u);
            }
            __syncwarp();
        }
        if (ptx::elect_one()) umma_commit2_mc(&done_bar, (uint16_t)3);
        __syncwarp();
    }
    ptx::mbar_wait(&done_bar, 0);         // arrives in BOTH CTAs when the pair's MMAs have retired
    ptx::tc_fence_after();
    if (mode == 1) {
        if (rank == 0 && warp == 1 && lane == 0) out[cluster] = (float)(clock64() - t0) / (float)(iters / 20 * 20);
    } else {
        float* dst = out + (((size_t)cluster * 2 + rank) * 128 + r) * PROBE_N;
#pragma unroll
        for (int j = 0; j < PROBE_N / 32; ++j) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(acc + ((uint32_t)(warp * 32) << 16) + j * 32, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[j * 32 + i] = __uint_as_float(v[i]);
        }
    }
    ptx::tc_fence_before();
    cluster_sync();                       // the peer's shared memory / TMEM stay alive until both are done
    if (warp == 0) tmem_dealloc2(tmem, 512);
}

}  // namespace rb
