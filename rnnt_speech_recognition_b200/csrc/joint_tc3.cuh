// joint_tc3.cuh -- third-generation fused joint kernel = joint_tc2 (z resident in TENSOR MEMORY, deep W ring,
// whole-stage MMA issue) + TMA-fed producers.
//
// v2's producers (coalesced global loads -> tanh -> smem staging -> 256-thread barrier -> re-read own row ->
// tcgen05.st) took ~2.1 k cycles per K block, and because z is single-buffered in TMEM that chain gates the first
// V-chunk of every tile (measured: 1.2 ms of a 3.8 ms kernel).  Here a dedicated warp TMA-loads the tile's
// pred rows (box [32 fp32 x UU rows], SWIZZLE_128B) and enc rows (box [64 x TT]) K block by K block into a
// ring, RUNNING AHEAD across tiles; each producer thread then reads ITS OWN lattice row straight from shared
// memory (conflict-free), applies tanh, packs to bf16 and writes its TMEM lane with tcgen05.st -- no staging
// pass, no inter-warp barrier, no register prefetch buffers.
//
// Roles (480 threads): warps 0-3 epilogue | 4-11 producers | 12 W TMA | 13 MMA | 14 enc/pred TMA
//
// MODE 0  forward: lse + (blank, label) log-probs per cell; nothing else leaves the SM
// MODE 1  recompute backward: same mainloop, epilogue writes the bf16 logit gradients, producers the bf16 z rows
// MODE 2  forward that KEEPS its activations (rnntb200JointDesc.keep_activations): MODE 0 + the fp16 softmax
//         numerators 2^(y - m) the epilogue computes anyway, the running maxima m, and the bf16 z rows;
//         dl_from_kept_kernel (below) then turns the numerators into logit gradients in one streaming pass.
#pragma once
#include <cuda_fp16.h>
#include "joint_tc2.cuh"

namespace rb {

constexpr int TC3_THREADS = 480;
constexpr int TC3_IN_STAGES = 2;

inline Tc2Geom tc3_geometry(int H, int V) {
    Tc2Geom g = tc2_geometry(H, V);
    if (!g.ok) return g;
    // smem: enc/pred ring (2 x (2 x 16 KB pred boxes + 4 KB enc box)) + W ring; W stages shrink to fit
    const size_t in_bytes = (size_t)TC3_IN_STAGES * (2 * 16384 + 4096);
    const size_t bias_bytes = V <= 4096 ? (size_t)V * 4 : 0;
    const size_t budget = 232448 - 1024 - 1024 - in_bytes - bias_bytes;
    g.stages = (int)(budget / ((size_t)g.ks * 8192));
    if (g.stages > 6) g.stages = 6;
    g.smem_bytes = 1024 + in_bytes + (size_t)g.stages * g.ks * 8192 + 1024 + bias_bytes;
    g.ok = g.stages >= 2;
    return g;
}

template <int MODE>
__global__ void __launch_bounds__(TC3_THREADS, 1) joint_tc3_kernel(const __grid_constant__ CUtensorMap tmap_wt,
                                                                   const __grid_constant__ CUtensorMap tmap_pred,
                                                                   const __grid_constant__ CUtensorMap tmap_enc,
                                                                   const JointTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int KB = p.KB, NCH = p.NCH, stages = p.stages, NBUF = p.nbuf, KS = p.ks;
    const uint32_t stage_bytes = (uint32_t)KS * 8192u;
    constexpr int NC = TC2_NC;
    uint8_t* insm = smem;                                 // TC3_IN_STAGES x {pred box k-half 0, k-half 1 (16 KB each, SW128), enc box (4 KB)}
    constexpr uint32_t IN_STAGE = 2 * 16384 + 4096;
    uint8_t* wsm = smem + TC3_IN_STAGES * IN_STAGE;       // stages x KS x [64 x 64] bf16, SW128 K-major (TMA)
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + (size_t)stages * stage_bytes);
    uint64_t* z_full = bars;                              // [TC_MAX_KB]      producers -> MMA (K block in TMEM)
    uint64_t* z_free = bars + TC_MAX_KB;                  //                  MMA -> producers
    uint64_t* w_full = z_free + 1;                        // [TC2_MAX_STAGES] TMA -> MMA
    uint64_t* w_empty = w_full + TC2_MAX_STAGES;          // [TC2_MAX_STAGES] MMA -> TMA
    uint64_t* acc_full = w_empty + TC2_MAX_STAGES;        // [TC2_MAX_NBUF]   MMA -> epilogue
    uint64_t* acc_empty = acc_full + TC2_MAX_NBUF;        // [TC2_MAX_NBUF]   epilogue -> MMA
    uint64_t* in_full = acc_empty + TC2_MAX_NBUF;         // [TC3_IN_STAGES]  input TMA -> producers
    uint64_t* in_empty = in_full + TC3_IN_STAGES;         // [TC3_IN_STAGES]  producers -> input TMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(in_empty + TC3_IN_STAGES);
    float* bias2 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 1024);   // bias * log2(e) for V <= 4096 (else read from global)
    const bool bias_in_smem = p.V <= 4096;
    if (bias_in_smem)
        for (int i = threadIdx.x; i < p.V; i += TC3_THREADS) bias2[i] = __ldg(p.bias + i) * 1.4426950408889634f;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_MAX_KB; ++i) ptx::mbar_init(&z_full[i], 8);
        ptx::mbar_init(z_free, 1);
        for (int i = 0; i < TC2_MAX_STAGES; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < TC2_MAX_NBUF; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 4); }
        for (int i = 0; i < TC3_IN_STAGES; ++i) { ptx::mbar_init(&in_full[i], 1); ptx::mbar_init(&in_empty[i], 8); }
        ptx::fence_barrier_init();
    }
    if (warp == 13) { ptx::tmem_alloc(tmem_ptr, TC_TMEM_COLS); ptx::tmem_relinquish(); }
    if (warp == 12 && lane == 0) ptx::prefetch_tmap(&tmap_wt);
    if (warp == 14 && lane == 0) { ptx::prefetch_tmap(&tmap_pred); ptx::prefetch_tmap(&tmap_enc); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t acc0 = tmem_base + (uint32_t)KB * 32;  // accumulator region starts after the z columns
    const int ntiles = p.nb * p.nTb * p.nUb;

    if (warp == 14) {
        // ===================== enc / pred TMA: one K block of the tile's rows per ring stage, runs ahead across tiles
        if (lane == 0 && !(p.dbg & 8)) {
            int st = 0; uint32_t ph = 0;
            const uint32_t tx = 2u * (uint32_t)p.UU * 128u + (uint32_t)p.TT * 256u;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const TileInfo ti = decode_tile(p, tile);
                if (!ti.valid) continue;
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait(&in_empty[st], ph ^ 1);
                    ptx::mbar_arrive_expect_tx(&in_full[st], tx);
                    uint8_t* base = insm + (size_t)st * IN_STAGE;
                    ptx::tma_load_2d(base, &tmap_pred, &in_full[st], kb * 64, ti.b * p.maxU + ti.u0);
                    ptx::tma_load_2d(base + 16384, &tmap_pred, &in_full[st], kb * 64 + 32, ti.b * p.maxU + ti.u0);
                    ptx::tma_load_2d(base + 32768, &tmap_enc, &in_full[st], kb * 64, ti.b * p.maxT + ti.t0);
                    if (++st == TC3_IN_STAGES) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 12) {
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
                if (!(p.dbg & 128)) {
                    ptx::mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
                    ptx::tc_fence_after();
                }
                const uint32_t d_tmem = acc0 + buf * NC;
                for (int kb0 = 0; kb0 < KB; kb0 += KS) {
                    if (!(p.dbg & 4)) ptx::mbar_wait(&w_full[stage], phase);
                    // TMA-written smem is consumed by the same (async) proxy the MMA reads through: the mbarrier wait
                    // alone orders it.  tcgen05.fence::after_thread_sync is only needed where OTHER THREADS' tcgen05
                    // traffic is involved: after acc_empty (epilogue tcgen05.ld) and z_full (producer tcgen05.st).
                    if (p.dbg & 64) ptx::tc_fence_after();
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
                        // steady state: the whole stage (up to 20 MMAs) AND its commits from ONE elected block with
                        // immediate operand offsets.  The single issuing thread is the pace-setter at N=64 (the probe
                        // reaches the 32-cycle floor only with >= 20 MMAs per block), so nothing else goes in between.
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
                            if (!(p.dbg & 4)) ptx::umma_commit(&w_empty[stage]);
                            if (kb0 + KS >= KB) ptx::umma_commit(&acc_full[buf]);
                        }
                        __syncwarp();
                        if (++stage == stages) { stage = 0; phase ^= 1; }
                        continue;
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
    } else if (warp >= 4 && warp < 12) {
        // ===================== producers (warps 4-11): thread = (lattice row r2 = TMEM lane, k-half hh) =====================
        const int pw = warp - 4, ptid = threadIdx.x - 128;
        const int q4 = pw & 3, hh = pw >> 2, r2 = q4 * 32 + lane;
        uint32_t it = 0; int st = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const TileInfo ti = decode_tile(p, tile);
            if (p.dbg & 8) continue;
            if (!ti.valid) {
                if (MODE != 0 && !p.slot) {  // uncompacted rows: the plain GEMMs reduce over ALL rows, padding tiles must read as zero
                    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
                    if (MODE == 1) {         // (MODE 2: dl_from_kept_kernel zero-fills the dlogits rows of padding tiles)
                        uint4* d4 = reinterpret_cast<uint4*>(p.dl + (size_t)tile * 128 * p.V);
                        for (int i = ptid; i < 128 * p.V / 8; i += 256) d4[i] = z4;
                    }
                    if (p.zb) {
                        const int h8 = p.H / 8 + 1;   // + the 8 columns at H (ones column): the dW GEMM reads H + 8 columns
                        for (int i = ptid; i < 128 * h8; i += 256)
                            *reinterpret_cast<uint4*>(p.zb + ((size_t)tile * 128 + i / h8) * p.zld + (i % h8) * 8) = z4;
                    }
                }
                continue;
            }
            const size_t rowbase = (size_t)(p.slot ? p.slot[tile] : tile) * 128;   // row block of this tile in dl / zb
            const int tl = r2 / p.UU, ul = r2 % p.UU;   // row of the enc box / of the pred box
            const bool ok = (ti.t0 + tl) < ti.Tn && (ti.u0 + ul) < ti.Un;
            for (int kb = 0; kb < KB; ++kb) {
                ptx::mbar_wait(&in_full[st], ph);
                const uint8_t* base = insm + (size_t)st * IN_STAGE;
                const uint8_t* prow = base + (size_t)hh * 16384 + ul * 128;          // SW128: chunk c at (c ^ (row & 7)) * 16
                const uint8_t* erow = base + 32768 + tl * 256 + hh * 128;            // plain layout
                uint32_t zr[16];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 q = *reinterpret_cast<const float4*>(prow + ((c ^ (ul & 7)) << 4));
                    const float4 e = *reinterpret_cast<const float4*>(erow + (c << 4));
                    if (ok) {   // (tanh.approx.bf16x2 / f16x2 lower to two scalar MUFU ops on sm_100a: no gain from packing)
                        zr[c * 2 + 0] = ptx::pack_bf16x2(ptx::tanh_approx(e.x + q.x), ptx::tanh_approx(e.y + q.y));
                        zr[c * 2 + 1] = ptx::pack_bf16x2(ptx::tanh_approx(e.z + q.z), ptx::tanh_approx(e.w + q.w));
                    } else {
                        zr[c * 2 + 0] = 0u; zr[c * 2 + 1] = 0u;
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&in_empty[st]);                       // this warp is done with the stage
                if (++st == TC3_IN_STAGES) { st = 0; ph ^= 1; }
                if (MODE != 0 && p.zb) {
                    __nv_bfloat16* zdst = p.zb + (rowbase + r2) * p.zld + kb * 64 + hh * 32;   // 64 bytes of this thread's row
                    ptx::st_global_256(zdst, zr);
                    ptx::st_global_256(zdst + 16, zr + 8);
                    // the ones column at index H that turns the dW GEMM's extra output row into db (written once per row,
                    // by the thread that owns the row's last 32 columns)
                    if (kb == KB - 1 && hh == 1)
                        *reinterpret_cast<uint4*>(p.zb + (rowbase + r2) * p.zld + p.H) = make_uint4(0x00003F80u, 0u, 0u, 0u);
                }
                if (kb == 0) ptx::mbar_wait(z_free, (it & 1) ^ 1);   // previous tile's MMAs have retired: z columns reusable
                ptx::tmem_st_32x16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(kb * 32 + hh * 16), zr);
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&z_full[kb]);
            }
            ++it;
        }
    } else if (warp < 4) {
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
                    if (p.dbg & 256) continue;
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(lane_addr + buf * NC + j * 32, v);
                    ptx::tmem_ld_wait();
                    if (p.dbg & 1) { s += __uint_as_float(v[0]); continue; }
                    const int col0 = c * NC + j * 32;
                    const float bv = bias_in_smem ? bias2[col0 + lane] : __ldg(p.bias + col0 + lane) * LOG2E;
                    float y[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        y[i] = fmaf(__uint_as_float(v[i]), LOG2E, __shfl_sync(0xffffffffu, bv, i));
                    if (MODE != 1) {
                        float gm = y[0];
#pragma unroll
                        for (int i = 1; i < 32; ++i) gm = fmaxf(gm, y[i]);
                        const float mn = fmaxf(m2, gm);
                        float acc = 0.f;
                        if (MODE == 0) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) acc += ptx::ex2_approx(y[i] - mn);
                        } else {
                            // keep the numerators: 2^(y - mn) in (0, 1] as fp16 (2^-11 relative), with mn beside them
                            uint32_t o[16];
#pragma unroll
                            for (int i = 0; i < 32; i += 2) {
                                const float e0 = ptx::ex2_approx(y[i] - mn), e1 = ptx::ex2_approx(y[i + 1] - mn);
                                acc += e0 + e1;
                                o[i >> 1] = ptx::pack_f16x2(e0, e1);
                            }
                            __nv_bfloat16* dst = p.dl + (rowbase + r) * p.V + col0;
                            ptx::st_global_256(dst, o);
                            ptx::st_global_256(dst + 16, o + 8);
                            p.gm[rowbase * (size_t)(p.V >> 5) + (size_t)(col0 >> 5) * 128 + r] = mn;   // [row block][group][row]: coalesced
                        }
                        s = s * ptx::ex2_approx(m2 - mn) + acc;
                        m2 = mn;
                        if (p.blank >= col0 && p.blank < col0 + 32) {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (col0 + i == p.blank) yb = y[i];
                        }
                        // logit[label_u]: the label differs per thread, so the wanted element sits at a DYNAMIC index of
                        // this thread's 32 registers: a binary select tree on the five index bits (31 selects, registers
                        // only -- the load/store pipe is the scarce unit of this kernel), run only when some lane of the
                        // warp has its label in this column group.
                        const int d = lab - col0;
                        const bool mine = (unsigned)d < 32u;
                        if (__any_sync(0xffffffffu, mine)) {
                            float s16[16], s8[8], s4[4];
#pragma unroll
                            for (int i = 0; i < 16; ++i) s16[i] = (d & 1) ? y[2 * i + 1] : y[2 * i];
#pragma unroll
                            for (int i = 0; i < 8; ++i) s8[i] = (d & 2) ? s16[2 * i + 1] : s16[2 * i];
#pragma unroll
                            for (int i = 0; i < 4; ++i) s4[i] = (d & 4) ? s8[2 * i + 1] : s8[2 * i];
                            const float s2a = (d & 8) ? s4[1] : s4[0], s2b = (d & 8) ? s4[3] : s4[2];
                            if (mine) yl = (d & 16) ? s2b : s2a;
                        }
                    } else {
                        uint32_t o[16];
#pragma unroll
                        for (int i = 0; i < 32; i += 2)
                            o[i >> 1] = ptx::pack_bf16x2(cg * ptx::ex2_approx(y[i] + kd2), cg * ptx::ex2_approx(y[i + 1] + kd2));
                        __nv_bfloat16* dst = p.dl + (rowbase + r) * p.V + col0;
                        ptx::st_global_256(dst, o);
                        ptx::st_global_256(dst + 16, o + 8);
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
            if (MODE != 1 && rv) {
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


// Backward of a forward that kept its activations (MODE 2): dlogits[row, v] = e[row, v] * g * 2^(gm[row, v/32] + kd)
// with e the kept fp16 numerators -- a pure streaming pass (2 bytes in, 2 bytes out per logit, IN PLACE: the bf16
// result overwrites the fp16 input) instead of a second projection on the tensor cores.  The two special columns
// (blank, label) receive their precomputed final values afterwards, from the thread that wrote that 16-byte vector.
//   grid = tiles of the launch (original order), block = 256: warp <-> row (16 rows each), lane <-> 16-byte vectors.
__device__ __forceinline__ void st_bf16_after(__nv_bfloat16* ptr, float val) {
    // a 2-byte store that the compiler may not move across the surrounding (differently typed) vector accesses;
    // same thread + same address, so the hardware keeps it after the 16-byte store it patches
    asm volatile("st.global.b16 [%0], %1;" :: "l"(ptr), "h"(__bfloat16_as_ushort(__float2bfloat16(val))) : "memory");
}
__global__ void __launch_bounds__(256) dl_from_kept_kernel(const JointTcParams p) {
    constexpr float LOG2E = 1.4426950408889634f;
    const int tile = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nvec = p.V >> 3;
    const TileInfo ti = decode_tile(p, tile);
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    if (!ti.valid) {
        if (!p.slot) {   // uncompacted rows: padding tiles must read as zero in the GEMMs
            uint4* d4 = reinterpret_cast<uint4*>(p.dl + (size_t)tile * 128 * p.V);
            for (int i = threadIdx.x; i < 128 * nvec; i += 256) d4[i] = z4;
        }
        return;
    }
    const size_t rowbase = (size_t)(p.slot ? p.slot[tile] : tile) * 128;
    // the tile's maxima [group][row] (coalesced in global) are staged as [row][group] (+1 word of padding per row:
    // conflict-free both ways) so that a row's lanes read consecutive words
    extern __shared__ float gm_s[];
    const int G = p.V >> 5;
    {
        const float* gsrc = p.gm + rowbase * (size_t)G;
        for (int i = threadIdx.x; i < G * 128; i += 256) gm_s[(i & 127) * (G + 1) + (i >> 7)] = gsrc[i];
    }
    __syncthreads();
    for (int r = warp; r < 128; r += 8) {
        const int t = ti.t0 + r / p.UU, u = ti.u0 + r % p.UU;
        const bool rv = t < ti.Tn && u < ti.Un;
        uint4* row4 = reinterpret_cast<uint4*>(p.dl + (rowbase + r) * p.V);
        if (!rv) {
            for (int v = lane; v < nvec; v += 32) row4[v] = z4;
            continue;
        }
        const float4 cf = p.coef[((long long)ti.b * p.maxT + t) * p.maxU + u];
        const float kd2 = cf.x * LOG2E, cg = cf.y;
        const int lab = (u < ti.Un - 1) ? p.labels[(size_t)ti.b * (p.maxU - 1) + u] : -1;
        const float* gmr = gm_s + r * (G + 1);
        for (int v0 = lane; v0 < nvec; v0 += 128) {
            uint4 x[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)                     // four independent 16-byte requests per lane before any use
                if (v0 + 32 * k < nvec) x[k] = row4[v0 + 32 * k];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int v = v0 + 32 * k;
                if (v < nvec) {
                    const float sc = cg * ptx::ex2_approx(gmr[v >> 2] + kd2);
                    const uint32_t w[4] = {x[k].x, x[k].y, x[k].z, x[k].w};
                    uint32_t o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float f0, f1;
                        ptx::unpack_f16x2(w[i], f0, f1);
                        o[i] = ptx::pack_bf16x2(f0 * sc, f1 * sc);
                    }
                    row4[v] = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        // the two special columns, patched by the lane that wrote their vector (vector index & 31 == lane)
        __nv_bfloat16* drow = reinterpret_cast<__nv_bfloat16*>(row4);
        if (((p.blank >> 3) & 31) == lane) st_bf16_after(drow + p.blank, cf.z);   // (label == blank: cf.w == cf.z)
        if (lab >= 0 && ((lab >> 3) & 31) == lane) st_bf16_after(drow + lab, cf.w);
    }
}

template <int MODE>
inline rnntStatus_t tc3_launch(const Tc2Geom& g3, const CUtensorMap& tm, const CUtensorMap& tmp, const CUtensorMap& tme,
                               const JointTcParams& p, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(joint_tc3_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)232448) != cudaSuccess)
            return RNNT_STATUS_EXECUTION_FAILED;
        attr_set = true;
    }
    const int ntiles = p.nb * p.nTb * p.nUb;
    const int grid = ntiles < tc_num_sms() ? ntiles : tc_num_sms();
    ScopedTimer tmr(MODE == 0 ? "joint_tc3_kernel<fwd>" : MODE == 1 ? "joint_tc3_kernel<dlogits>" : "joint_tc3_kernel<fwd+keep>", s);
    joint_tc3_kernel<MODE><<<grid, TC3_THREADS, g3.smem_bytes, s>>>(tm, tmp, tme, p);
    return cudaGetLastError() == cudaSuccess ? RNNT_STATUS_SUCCESS : RNNT_STATUS_EXECUTION_FAILED;
}

}  // namespace rb
