// joint_tc2.cuh -- second-generation fused joint kernel: the A operand (z = tanh(enc+pred), bf16) lives in
// TENSOR MEMORY instead of shared memory.
//
// Why: with z resident in smem (v1, joint_tc.cuh) only two 32 KB W^T stages fit beside it, and the W stream
// (64 B/clk/SM at full MMA rate) becomes latency-bound -- ncu on v1: 63 k cycles per 128-cell tile against a
// 20.5 k-cycle MMA floor, tensor pipe 33 % active.  Moving z to TMEM (H/2 = 320 of the 512 columns) frees
// ~190 KB of smem for a 24-stage W ring and lets tcgen05.mma read A at TMEM bandwidth.
//
// Roles (448 threads, 1 CTA/SM, persistent over 128-cell tiles):
//   warps 0-3   epilogue: tcgen05.ld of 64-column accumulator chunks (3 TMEM buffers), online LSE / dlogits
//   warps 4-11  producers: (1) coalesced enc/pred loads -> tanh -> bf16 -> smem staging tile (SW128 pattern,
//               conflict-free), (2) after a 256-thread named barrier each thread re-reads ITS OWN row
//               (thread = TMEM lane) and writes it to TMEM with tcgen05.st.32x32b.x16
//   warp 12     TMA: W^T boxes [64 v x 64 k] (8 KB) through the w_full/w_empty ring
//   warp 13     MMA: tcgen05.mma.kind::f16 with A from TMEM, B from smem (M=128, N=64, K=16)
#pragma once
#include "joint_tc.cuh"

namespace rb {

constexpr int TC2_THREADS = 448;
constexpr int TC2_NC = 64;
constexpr int TC2_MAX_STAGES = 24;
constexpr int TC2_MAX_NBUF = 4;

struct Tc2Geom { int nbuf, stages, zcols, ks; size_t smem_bytes; bool ok; };
inline Tc2Geom tc2_geometry(int H, int V) {
    Tc2Geom g{};
    if (H % 64 || V % 64) return g;
    g.zcols = (H / 64) * 32;
    const int acc_cols = TC_TMEM_COLS - g.zcols;
    g.nbuf = acc_cols / TC2_NC;
    if (g.nbuf > TC2_MAX_NBUF) g.nbuf = TC2_MAX_NBUF;
    if (g.nbuf < 2) return g;
    // One W stage = ks K-blocks ([64 v x 64 k] boxes, 8 KB each): the MMA thread pays one mbarrier wait and one
    // commit per 4*ks MMAs.  (With ks = 1 the single issuing thread, not the tensor pipe, set the pace: 123 k
    // cycles per tile measured on B200.)
    const int KB = H / 64;
    g.ks = 1;
    for (int k = 5; k >= 1; --k)
        if (KB % k == 0) { g.ks = k; break; }   // largest divisor of KB that is <= 5: stages are never partial
    g.stages = (int)((size_t)(20 * 8192) / ((size_t)g.ks * 8192));
    if (g.stages > TC2_MAX_STAGES) g.stages = TC2_MAX_STAGES;
    g.smem_bytes = 1024 /*align*/ + 2 * 16384 + (size_t)g.stages * g.ks * 8192 + 1024 /*barriers*/;
    g.ok = g.smem_bytes <= 232448;
    return g;
}

template <int MODE>
__global__ void __launch_bounds__(TC2_THREADS, 1) joint_tc2_kernel(const __grid_constant__ CUtensorMap tmap_wt,
                                                                   const JointTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int KB = p.KB, NCH = p.NCH, stages = p.stages, NBUF = p.nbuf, KS = p.ks;
    const uint32_t stage_bytes = (uint32_t)KS * 8192u;
    constexpr int NC = TC2_NC;
    uint8_t* sb = smem;                                   // 2 x [128 x 64] bf16 staging tiles (SW128 pattern)
    uint8_t* wsm = smem + 2 * 16384;                      // stages x KS x [64 x 64] bf16, SW128 K-major (TMA)
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + (size_t)stages * stage_bytes);
    uint64_t* z_full = bars;                              // [TC_MAX_KB]      producers -> MMA (K block in TMEM)
    uint64_t* z_free = bars + TC_MAX_KB;                  //                  MMA -> producers
    uint64_t* w_full = z_free + 1;                        // [TC2_MAX_STAGES] TMA -> MMA
    uint64_t* w_empty = w_full + TC2_MAX_STAGES;          // [TC2_MAX_STAGES] MMA -> TMA
    uint64_t* acc_full = w_empty + TC2_MAX_STAGES;        // [TC2_MAX_NBUF]   MMA -> epilogue
    uint64_t* acc_empty = acc_full + TC2_MAX_NBUF;        // [TC2_MAX_NBUF]   epilogue -> MMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + TC2_MAX_NBUF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_MAX_KB; ++i) ptx::mbar_init(&z_full[i], 8);
        ptx::mbar_init(z_free, 1);
        for (int i = 0; i < TC2_MAX_STAGES; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < TC2_MAX_NBUF; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 4); }
        ptx::fence_barrier_init();
    }
    if (warp == 13) { ptx::tmem_alloc(tmem_ptr, TC_TMEM_COLS); ptx::tmem_relinquish(); }
    if (warp == 12 && lane == 0) ptx::prefetch_tmap(&tmap_wt);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t acc0 = tmem_base + (uint32_t)KB * 32;  // accumulator region starts after the z columns
    const int ntiles = p.nb * p.nTb * p.nUb;

    if (warp == 12) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                if (!decode_tile(p, tile).valid) continue;
                if (p.dbg & 4) continue;
                for (int c = 0; c < NCH; ++c)
                    for (int kb0 = 0; kb0 < KB; kb0 += KS) {   // KB % KS == 0: one 3-D box = KS K-block slabs
                        ptx::mbar_wait(&w_empty[stage], phase ^ 1);
                        ptx::mbar_arrive_expect_tx(&w_full[stage], stage_bytes);
                        ptx::tma_load_3d(wsm + (size_t)stage * stage_bytes, &tmap_wt, &w_full[stage], 0, c * NC, kb0);
                        if (++stage == stages) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 13) {
        // ===================== MMA issuer: A from TMEM, B from smem =====================
        // The WHOLE warp runs this loop convergently and one elected lane issues: descriptors, TMEM addresses and
        // barrier addresses then stay in uniform registers.  (Issuing from inside `if (lane == 0)` made every
        // operand a vector register that had to be moved to the uniform datapath per instruction -- measured
        // 117 cycles per N=64 MMA instead of the 32-cycle dispatch floor.)
        const uint32_t idesc = ptx::umma_idesc_bf16(128, NC);
        int stage = 0; uint32_t phase = 0, g = 0, it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            if (!decode_tile(p, tile).valid) continue;
            for (int c = 0; c < NCH; ++c, ++g) {
                const uint32_t buf = g % NBUF, use = g / NBUF;
                ptx::mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = acc0 + buf * NC;
                for (int kb0 = 0; kb0 < KB; kb0 += KS) {
                    if (!(p.dbg & 4)) ptx::mbar_wait(&w_full[stage], phase);
                    ptx::tc_fence_after();   // once per stage (4*KS MMAs)
                    const uint64_t bdesc0 = ptx::umma_desc_k_sw128(ptx::smem_u32(wsm + (size_t)stage * stage_bytes));
                    const uint32_t a_st = tmem_base + (uint32_t)kb0 * 32;
                    if (c == 0 && !(p.dbg & 8)) {
                        // first chunk of a tile: each K block of z must have landed in TMEM before it is read
                        for (int i = 0; i < KS; ++i) {
                            ptx::mbar_wait(&z_full[kb0 + i], it & 1);
                            ptx::tc_fence_after();
                            if (ptx::elect_one()) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    ptx::umma_bf16_ts(d_tmem, a_st + i * 32 + k * 8, bdesc0 + (uint64_t)(i * 512 + k * 2),
                                                      idesc, (uint32_t)((kb0 | i | k) != 0));
                            }
                            __syncwarp();
                        }
                    } else {
                        // steady state: the whole stage (up to 20 MMAs) from ONE elected block with immediate
                        // operand offsets.  A per-K-block loop cost ~40 cycles of issue overhead per 46-cycle MMA.
                        if (ptx::elect_one()) {
#pragma unroll
                            for (int i = 0; i < 5; ++i) {
                                if (i < KS) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        ptx::umma_bf16_ts(d_tmem, a_st + i * 32 + k * 8,
                                                          bdesc0 + (uint64_t)(i * 512 + k * 2), idesc,
                                                          (i | k) ? 1u : (uint32_t)(kb0 != 0));
                                }
                            }
                        }
                        __syncwarp();
                    }
                    if (ptx::elect_one()) {
                        if (!(p.dbg & 4)) ptx::umma_commit(&w_empty[stage]);
                        if (kb0 + KS >= KB) ptx::umma_commit(&acc_full[buf]);
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
            if (!(p.dbg & 8) && ptx::elect_one()) ptx::umma_commit(z_free);
            __syncwarp();
            ++it;
        }
    } else if (warp >= 4) {
        // ===================== producers (warps 4-11, 256 threads) =====================
        const int pw = warp - 4, ptid = threadIdx.x - 128;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const TileInfo ti = decode_tile(p, tile);
            if (p.dbg & 8) continue;
            if (!ti.valid) {
                if (MODE == 1 && !p.slot) {  // uncompacted rows: the plain GEMMs reduce over ALL rows, padding tiles must read as zero
                    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
                    uint4* d4 = reinterpret_cast<uint4*>(p.dl + (size_t)tile * 128 * p.V);
                    for (int i = ptid; i < 128 * p.V / 8; i += 256) d4[i] = z4;
                    if (p.zb) {
                        const int h8 = p.H / 8;
                        for (int i = ptid; i < 128 * h8; i += 256)
                            *reinterpret_cast<uint4*>(p.zb + ((size_t)tile * 128 + i / h8) * p.zld + (i % h8) * 8) = z4;
                    }
                }
                continue;
            }
            const size_t rowbase = (size_t)(p.slot ? p.slot[tile] : tile) * 128;   // row block of this tile in dl / zb
            uint32_t eo[4], qo[4], soff[4];
            bool ok[4];
            const float4* enc4 = reinterpret_cast<const float4*>(p.enc);
            const float4* pred4 = reinterpret_cast<const float4*>(p.pred);
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
                const int r = pass * 32 + pw * 4 + (lane >> 3), ch = lane & 7;
                const int t = ti.t0 + r / p.UU, u = ti.u0 + r % p.UU;
                ok[pass] = t < ti.Tn && u < ti.Un;
                eo[pass] = (uint32_t)((((size_t)ti.b * p.maxT + (ok[pass] ? t : 0)) * p.H + ch * 8) >> 2);
                qo[pass] = (uint32_t)((((size_t)ti.b * p.maxU + (ok[pass] ? u : 0)) * p.H + ch * 8) >> 2);
                soff[pass] = r * 128 + ((ch ^ (r & 7)) << 4);
            }
            // pred rows (distinct per lane group, L2 latency) are prefetched one K block ahead into registers;
            // enc rows (shared by the whole tile when TT == 1, L1-resident) are loaded just in time.
            float4 bufA[8], bufB[8];
            auto issue = [&](int kb, float4* buf) {
#pragma unroll
                for (int pass = 0; pass < 4; ++pass) {
                    const float4* q = pred4 + qo[pass] + kb * 16;
                    buf[pass * 2 + 0] = __ldg(q); buf[pass * 2 + 1] = __ldg(q + 1);
                }
            };
            const int q4 = pw & 3, hh = pw >> 2, r2 = q4 * 32 + lane;   // phase 2: this thread owns TMEM lane r2
            auto produce = [&](int kb, const float4* buf) {
                uint8_t* stg = sb + (size_t)(kb & 1) * 16384;
                if (kb + 1 < KB) {   // pull the next K block's enc segment into L1 while this block's tanh work runs
#pragma unroll
                    for (int pass = 0; pass < 4; ++pass)
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(enc4 + eo[pass] + (kb + 1) * 16));
                }
                // phase 1: tanh -> bf16 -> staging tile (8 lanes cover one row's 64 k: conflict-free 16-byte stores)
#pragma unroll
                for (int pass = 0; pass < 4; ++pass) {
                    const float4* e = enc4 + eo[pass] + kb * 16;
                    const float4 e0 = __ldg(e), e1 = __ldg(e + 1), q0 = buf[pass * 2], q1 = buf[pass * 2 + 1];
                    uint4 packed = make_uint4(0u, 0u, 0u, 0u);
                    if (ok[pass] && (p.dbg & 2)) {
                        packed.x = ptx::pack_bf16x2(e0.x + q0.x, e0.y + q0.y); packed.y = ptx::pack_bf16x2(e0.z + q0.z, e0.w + q0.w);
                        packed.z = ptx::pack_bf16x2(e1.x + q1.x, e1.y + q1.y); packed.w = ptx::pack_bf16x2(e1.z + q1.z, e1.w + q1.w);
                    } else if (ok[pass]) {
                        packed.x = ptx::pack_bf16x2(ptx::tanh_approx(e0.x + q0.x), ptx::tanh_approx(e0.y + q0.y));
                        packed.y = ptx::pack_bf16x2(ptx::tanh_approx(e0.z + q0.z), ptx::tanh_approx(e0.w + q0.w));
                        packed.z = ptx::pack_bf16x2(ptx::tanh_approx(e1.x + q1.x), ptx::tanh_approx(e1.y + q1.y));
                        packed.w = ptx::pack_bf16x2(ptx::tanh_approx(e1.z + q1.z), ptx::tanh_approx(e1.w + q1.w));
                    }
                    *reinterpret_cast<uint4*>(stg + soff[pass]) = packed;
                    if (MODE == 1 && p.zb) {
                        const int r = pass * 32 + pw * 4 + (lane >> 3);
                        *reinterpret_cast<uint4*>(p.zb + (rowbase + r) * p.zld + kb * 64 + (lane & 7) * 8) = packed;
                    }
                }
                ptx::named_bar_sync(1, 256);
                if (kb == 0) ptx::mbar_wait(z_free, (it & 1) ^ 1);   // previous tile's MMAs have retired: z columns reusable
                // phase 2: own row, chunks hh*4 .. hh*4+3 (32 k = 16 TMEM columns), un-swizzled
                uint32_t zr[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = hh * 4 + j;
                    const uint4 v = *reinterpret_cast<const uint4*>(stg + r2 * 128 + ((c ^ (r2 & 7)) << 4));
                    zr[j * 4 + 0] = v.x; zr[j * 4 + 1] = v.y; zr[j * 4 + 2] = v.z; zr[j * 4 + 3] = v.w;
                }
                if (p.swap) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) zr[j] = __byte_perm(zr[j], 0, 0x1032);
                }
                ptx::tmem_st_32x16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(kb * 32 + hh * 16), zr);
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&z_full[kb]);
            };
            issue(0, bufA);
            for (int kb = 0; kb < KB; kb += 2) {
                if (kb + 1 < KB) issue(kb + 1, bufB);
                produce(kb, bufA);
                if (kb + 1 < KB) {
                    if (kb + 2 < KB) issue(kb + 2, bufA);
                    produce(kb + 1, bufB);
                }
            }
            ++it;
        }
    } else {
        // ===================== epilogue warps 0-3: thread = lattice cell (TMEM lane) =====================
        constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const TileInfo ti = decode_tile(p, tile);
            if (!ti.valid) continue;
            const size_t rowbase = (size_t)(p.slot ? p.slot[tile] : tile) * 128;   // row block of this tile in dl / zb
            const int r = warp * 32 + lane;
            const int t = ti.t0 + r / p.UU, u = ti.u0 + r % p.UU;
            const bool rv = t < ti.Tn && u < ti.Un;
            const int lab = (rv && u < ti.Un - 1) ? p.labels[(size_t)ti.b * (p.maxU - 1) + u] : -1;
            const long long cell = ((long long)ti.b * p.maxT + t) * p.maxU + u;
            float m2 = -CUDART_INF_F, s = 0.f, yb = 0.f, yl = 0.f;
            float kd2 = -CUDART_INF_F, cg = 0.f, csb = 0.f, csl = 0.f;
            if (MODE == 1 && rv) {
                const float4 cf = p.coef[cell];
                kd2 = cf.x * LOG2E; cg = cf.y; csb = cf.z; csl = cf.w;
            }
            const uint32_t lane_addr = acc0 + ((uint32_t)(warp * 32) << 16);
            for (int c = 0; c < NCH; ++c, ++g) {
                const uint32_t buf = g % NBUF, use = g / NBUF;
                ptx::mbar_wait(&acc_full[buf], use & 1);
                ptx::tc_fence_after();
#pragma unroll
                for (int j = 0; j < NC / 32; ++j) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(lane_addr + buf * NC + j * 32, v);
                    ptx::tmem_ld_wait();
                    if (p.dbg & 1) { s += __uint_as_float(v[0]); continue; }
                    const int col0 = c * NC + j * 32;
                    const float bv = __ldg(p.bias + col0 + lane) * LOG2E;
                    float y[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        y[i] = fmaf(__uint_as_float(v[i]), LOG2E, __shfl_sync(0xffffffffu, bv, i));
                    if (MODE == 0) {
                        float gm = y[0];
#pragma unroll
                        for (int i = 1; i < 32; ++i) gm = fmaxf(gm, y[i]);
                        const float mn = fmaxf(m2, gm);
                        float acc = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc += ptx::ex2_approx(y[i] - mn);
                        s = s * ptx::ex2_approx(m2 - mn) + acc;
                        m2 = mn;
                        if (p.blank >= col0 && p.blank < col0 + 32) {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (col0 + i == p.blank) yb = y[i];
                        }
                        const int d = lab - col0;
#pragma unroll
                        for (int i = 0; i < 32; ++i) yl = (i == d) ? y[i] : yl;
                    } else {
                        uint32_t o[16];
#pragma unroll
                        for (int i = 0; i < 32; i += 2)
                            o[i >> 1] = ptx::pack_bf16x2(cg * ptx::ex2_approx(y[i] + kd2), cg * ptx::ex2_approx(y[i + 1] + kd2));
                        uint4* dst = reinterpret_cast<uint4*>(p.dl + (rowbase + r) * p.V + col0);
                        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
                        dst[2] = make_uint4(o[8], o[9], o[10], o[11]);
                        dst[3] = make_uint4(o[12], o[13], o[14], o[15]);
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
            }
            if (MODE == 1 && rv) {   // the two special columns: final values precomputed by cell_coef_kernel
                __nv_bfloat16* drow = p.dl + (rowbase + r) * p.V;
                drow[p.blank] = __float2bfloat16(csb);
                if (lab >= 0) drow[lab] = __float2bfloat16(csl);
            }
            if (MODE == 0 && rv) {
                const float lse2 = m2 + log2f(s);
                p.lse[cell] = lse2 * LN2;
                const long long k = sk_index(ti.b, t, u, p.maxU, p.SK);
                p.lpb[k] = (yb - lse2) * LN2;
                if (u < ti.Un - 1) p.lpl[k] = (yl - lse2) * LN2;
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 13) ptx::tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

template <int MODE>
inline rnntStatus_t tc2_launch(const Tc2Geom& g2, const CUtensorMap& tm, const JointTcParams& p, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(joint_tc2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)232448) != cudaSuccess)
            return RNNT_STATUS_EXECUTION_FAILED;
        attr_set = true;
    }
    const int ntiles = p.nb * p.nTb * p.nUb;
    const int grid = ntiles < tc_num_sms() ? ntiles : tc_num_sms();
    ScopedTimer tmr(MODE == 0 ? "joint_tc2_kernel<fwd>" : "joint_tc2_kernel<dlogits>", s);
    joint_tc2_kernel<MODE><<<grid, TC2_THREADS, g2.smem_bytes, s>>>(tm, p);
    return cudaGetLastError() == cudaSuccess ? RNNT_STATUS_SUCCESS : RNNT_STATUS_EXECUTION_FAILED;
}

}  // namespace rb
