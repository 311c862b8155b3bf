// bwd_tc.cuh -- the two backward contractions of the joint (TF autograd through model.py:162-166,
// triggered at run_rnnt.py:284) as hand-written tcgen05 kernels that consume what the forward KEPT (fp16 softmax
// numerators e = 2^(y - m) and the running maxima m per 32-column group) and fuse their neighbours away:
//
//   bwd_dz_kernel   dZ[rows,H] = dl[rows,V] . W^T   with  dl = e * g*2^(m + kd)  formed in shared memory by a SIMT
//                   prologue on the TMA-loaded A stage (the two special columns patched from cell_coef_kernel's
//                   values), and an epilogue that applies (1 - tanh^2(enc+pred)), sums the tile over u (-> d_enc,
//                   accumulated in REGISTERS along a run of tiles and stored once) and over t (-> one fp32 partial
//                   plane of d_pred per t-block).  Neither dl nor dZ ever exists in HBM.
//   bwd_dw_kernel   dW[H,V] (+)= z^T . dl   as a split-K GEMM over the lattice rows: the A operand z^T is REGENERATED
//                   from enc/pred by producer warps (K-major SWIZZLE_128B tiles), the B operand is the kept
//                   numerators loaded MN-major by TMA and scaled in place; db = column sums of dl fall out of the
//                   scaling pass.  Partial tiles go to fp32 planes (deterministic), summed by sum_planes_kernel.
//
// Tile geometry: 16 x 8 lattice tiles (TT = 16 time steps, UU = 8 label positions), row r = tl*8 + ul of a tile is
// TMEM lane r; rows of tile `tile` live at row block slot[tile] of the kept arrays.
#pragma once
#include "joint_tc.cuh"

namespace rb {

constexpr int BW_TT = 16, BW_UU = 8;
constexpr int DZ_THREADS = 448;      // warp 0 TMA | 1 MMA | 2-5 A-stage scalers | 6-13 epilogue
constexpr int DZ_STAGES = 3;
constexpr int DW_THREADS = 448;      // warp 0 TMA | 1 MMA | 2-5 B-stage scalers | 6-13 z producers, then epilogue
constexpr int DW_STAGES = 3;
constexpr int DW_NV = 256;           // vocabulary columns per dW output tile

struct BwdParams {
    const float* enc; const float* pred;
    const int* labels; const int* xlen; const int* ylen;
    int maxT, maxU, H, V, blank;
    int nTb, nUb, b0, nb;
    const int* slot;            // tile -> row block of the kept arrays (-1: tile outside the valid lattice)
    const int* tile_of_slot;    // inverse map (valid tiles only)
    const int* count;           // number of valid tiles (device word)
    const float4* coef;         // per cell (kd, g, dl_blank, dl_label), natural cell order
    const float* gm;            // [row block][V/32][128] running maxima of the kept numerators (log2 domain)
    // ---- dZ kernel
    int NP, NCZ, priv, sh, odd_base;   // passes over H, columns per pass, private / shared accumulator columns
    float* d_enc;               // (B, maxT, H), rows of this launch's utterances are fully written
    float* ppred;               // (nTb, nb, maxU, H) partial planes of d_pred
    // ---- dW kernel
    int nVT, nD, nS, S_d, S_s, Hrows;  // v-tiles, double / single h-items per v-tile, their split counts, rows per plane
    float* dWp;                 // (S_max, Hrows, V) partial planes of dW
    float* dbp;                 // (S_d, V) partial planes of db
    int accumulate;             // planes already hold earlier utterance chunks
};

// ---------------------------------------------------------------------------------------------------------------
// shared helpers
// ---------------------------------------------------------------------------------------------------------------
// 8 kept numerators (fp16) * scale -> 8 bf16, with up to two entries replaced by precomputed finals
__device__ __forceinline__ uint4 scale_chunk(const uint4 x, float s, int ib, float vb, int il, float vl, float* colsum) {
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
    float f[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float lo, hi;
        ptx::unpack_f16x2(w[i], lo, hi);
        f[2 * i] = lo * s; f[2 * i + 1] = hi * s;
    }
    if ((unsigned)ib < 8u) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = (i == ib) ? vb : f[i];
    }
    if ((unsigned)il < 8u) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = (i == il) ? vl : f[i];
    }
    if (colsum) {
#pragma unroll
        for (int i = 0; i < 8; ++i) colsum[i] += f[i];
    }
    return make_uint4(ptx::pack_bf16x2(f[0], f[1]), ptx::pack_bf16x2(f[2], f[3]), ptx::pack_bf16x2(f[4], f[5]),
                      ptx::pack_bf16x2(f[6], f[7]));
}

// ---------------------------------------------------------------------------------------------------------------
// dZ kernel
// ---------------------------------------------------------------------------------------------------------------
// Work order (identical in every role): runs = (utterance, t-block) assigned round-robin to CTAs; inside a run the
// valid tiles along u; inside a tile NP passes over H.  Accumulators: an even unit uses TMEM columns [0, NCZ), an odd
// unit [priv .. NCZ) (the SHARED zone, drained first by the epilogue) + [odd_base, 512): two units are in flight
// (MMA of unit q+1 over the epilogue of unit q) although 2*NCZ may exceed the 512 columns.
struct DzSmem {
    static constexpr int kBars = 64;
};

__global__ void __launch_bounds__(DZ_THREADS, 1) bwd_dz_kernel(const __grid_constant__ CUtensorMap tmap_e,
                                                               const __grid_constant__ CUtensorMap tmap_wp,
                                                               const __grid_constant__ CUtensorMap tmap_ws,
                                                               const BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int NCZ = p.NCZ, NP = p.NP, priv = p.priv, sh = p.sh, KBV = p.V >> 6, G = p.V >> 5;
    const uint32_t stage_bytes = 16384u + (uint32_t)NCZ * 128u;
    uint8_t* dpb = smem + (size_t)DZ_STAGES * stage_bytes;                 // [hh][buf][4][8][36] floats
    constexpr int DP_ONE = 4 * 8 * 36;                                     // floats per (hh, buf)
    uint64_t* bars = reinterpret_cast<uint64_t*>(dpb + 2 * 2 * DP_ONE * 4);
    uint64_t* stage_full = bars;                     // [DZ_STAGES] TMA -> scalers, MMA
    uint64_t* a_ready = stage_full + DZ_STAGES;      // [DZ_STAGES] scalers -> MMA
    uint64_t* stage_empty = a_ready + DZ_STAGES;     // [DZ_STAGES] MMA -> TMA
    uint64_t* acc_full = stage_empty + DZ_STAGES;    // [2] MMA -> epilogue (unit parity)
    uint64_t* priv_free = acc_full + 2;              // [2] epilogue -> MMA
    uint64_t* shared_free = priv_free + 2;           // [1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(shared_free + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < DZ_STAGES; ++i) { ptx::mbar_init(&stage_full[i], 1); ptx::mbar_init(&a_ready[i], 4); ptx::mbar_init(&stage_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&priv_free[i], 8); }
        ptx::mbar_init(shared_free, 8);
        ptx::fence_barrier_init();
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_ptr, TC_TMEM_COLS); ptx::tmem_relinquish(); }
    if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tmap_e); ptx::prefetch_tmap(&tmap_wp); ptx::prefetch_tmap(&tmap_ws); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int nruns = p.nb * p.nTb;

    if (warp == 0) {
        // ===================== TMA: A = kept numerators [128 rows x 64 v], B = W rows [NCZ h x 64 v] =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int run = blockIdx.x; run < nruns; run += gridDim.x) {
                const int bl = run / p.nTb, tb = run - bl * p.nTb, b = p.b0 + bl;
                const int Tn = p.xlen[b], Un = p.ylen[b] + 1;
                if (tb * BW_TT >= Tn) continue;
                const int nub = (Un + BW_UU - 1) / BW_UU;
                for (int ub = 0; ub < nub; ++ub) {
                    const int tile = (bl * p.nTb + tb) * p.nUb + ub;
                    const int sl = p.slot ? p.slot[tile] : tile;
                    for (int pass = 0; pass < NP; ++pass) {
                        const int n0 = pass * NCZ;
                        for (int kb = 0; kb < KBV; ++kb) {
                            ptx::mbar_wait(&stage_empty[stage], phase ^ 1);
                            ptx::mbar_arrive_expect_tx(&stage_full[stage], stage_bytes);
                            uint8_t* st = smem + (size_t)stage * stage_bytes;
                            ptx::tma_load_2d(st, &tmap_e, &stage_full[stage], kb * 64, sl * 128);
                            ptx::tma_load_2d(st + 16384, &tmap_wp, &stage_full[stage], kb * 64, n0);
                            if (sh) ptx::tma_load_2d(st + 16384 + (size_t)priv * 128, &tmap_ws, &stage_full[stage], kb * 64, n0 + priv);
                            if (++stage == DZ_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA: D[128 x NCZ] += A[128 x 64] . B[NCZ x 64]^T, both operands K-major in smem =====================
        const uint32_t idescP = ptx::umma_idesc_bf16(128, priv), idescS = ptx::umma_idesc_bf16(128, sh ? sh : 16);
        int stage = 0; uint32_t phase = 0, q = 0;
        for (int run = blockIdx.x; run < nruns; run += gridDim.x) {
            const int bl = run / p.nTb, tb = run - bl * p.nTb, b = p.b0 + bl;
            const int Tn = p.xlen[b], Un = p.ylen[b] + 1;
            if (tb * BW_TT >= Tn) continue;
            const int nub = (Un + BW_UU - 1) / BW_UU;
            for (int ub = 0; ub < nub; ++ub)
                for (int pass = 0; pass < NP; ++pass, ++q) {
                    const uint32_t par = q & 1;
                    ptx::mbar_wait(&priv_free[par], ((q >> 1) & 1) ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t dP = tmem_base + (par ? (uint32_t)p.odd_base : 0u), dS = tmem_base + (uint32_t)priv;
                    for (int kb = 0; kb < KBV; ++kb) {
                        ptx::mbar_wait(&stage_full[stage], phase);
                        ptx::mbar_wait(&a_ready[stage], phase);
                        ptx::tc_fence_after();
                        const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * stage_bytes);
                        const uint64_t ad = ptx::umma_desc_k_sw128(sa), bp = ptx::umma_desc_k_sw128(sa + 16384u),
                                       bs = ptx::umma_desc_k_sw128(sa + 16384u + (uint32_t)priv * 128u);
                        if (ptx::elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                ptx::umma_bf16(dP, ad + (uint64_t)(k * 2), bp + (uint64_t)(k * 2), idescP, (uint32_t)((kb | k) != 0));
                        }
                        __syncwarp();
                        if (sh) {
                            if (kb == 0) {   // the shared zone still holds the previous unit until its epilogue has drained it
                                ptx::mbar_wait(shared_free, (q & 1) ^ 1);
                                ptx::tc_fence_after();
                            }
                            if (ptx::elect_one()) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    ptx::umma_bf16(dS, ad + (uint64_t)(k * 2), bs + (uint64_t)(k * 2), idescS, (uint32_t)((kb | k) != 0));
                            }
                            __syncwarp();
                        }
                        if (ptx::elect_one()) {
                            ptx::umma_commit(&stage_empty[stage]);
                            if (kb == KBV - 1) ptx::umma_commit(&acc_full[par]);
                        }
                        __syncwarp();
                        if (++stage == DZ_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
        }
    } else if (warp < 6) {
        // ===================== A-stage scalers (warps 2-5): thread = lattice row of the tile =====================
        constexpr float LOG2E = 1.4426950408889634f;
        const int r = threadIdx.x - 64;
        int stage = 0; uint32_t phase = 0;
        for (int run = blockIdx.x; run < nruns; run += gridDim.x) {
            const int bl = run / p.nTb, tb = run - bl * p.nTb, b = p.b0 + bl;
            const int Tn = p.xlen[b], Un = p.ylen[b] + 1;
            if (tb * BW_TT >= Tn) continue;
            const int nub = (Un + BW_UU - 1) / BW_UU;
            const int t = tb * BW_TT + (r >> 3);
            for (int ub = 0; ub < nub; ++ub) {
                const int tile = (bl * p.nTb + tb) * p.nUb + ub;
                const int sl = p.slot ? p.slot[tile] : tile;
                const int u = ub * BW_UU + (r & 7);
                const bool rv = t < Tn && u < Un;
                float4 cf = make_float4(-CUDART_INF_F, 0.f, 0.f, 0.f);
                int lab = -1;
                if (rv) {
                    cf = p.coef[((long long)b * p.maxT + t) * p.maxU + u];
                    if (u < Un - 1) lab = p.labels[(size_t)b * (p.maxU - 1) + u];
                }
                const float kd2 = cf.x * LOG2E;
                const float* gmr = p.gm + (size_t)sl * G * 128 + r;
                for (int pass = 0; pass < NP; ++pass) {
                    float g0 = gmr[0], g1 = gmr[128];
                    for (int kb = 0; kb < KBV; ++kb) {
                        const float s0 = rv ? cf.y * ptx::ex2_approx(g0 + kd2) : 0.f, s1 = rv ? cf.y * ptx::ex2_approx(g1 + kd2) : 0.f;
                        if (kb + 1 < KBV) { g0 = gmr[(size_t)(2 * kb + 2) * 128]; g1 = gmr[(size_t)(2 * kb + 3) * 128]; }
                        const int db = p.blank - kb * 64, dl = lab - kb * 64;   // position of the special columns inside this K block
                        ptx::mbar_wait(&stage_full[stage], phase);
                        uint8_t* row = smem + (size_t)stage * stage_bytes + r * 128;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            uint4* ptr = reinterpret_cast<uint4*>(row + ((c ^ (r & 7)) << 4));
                            const uint4 x = *ptr;
                            *ptr = scale_chunk(x, c < 4 ? s0 : s1, rv ? db - c * 8 : -1, cf.z, rv ? dl - c * 8 : -1, cf.w, nullptr);
                        }
                        ptx::fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&a_ready[stage]);
                        if (++stage == DZ_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else {
        // ===================== epilogue (warps 6-13): g = acc * (1 - tanh^2), tile sums =====================
        const int qd = warp & 3, hh = (warp - 6) >> 2, qslot = (warp - 6) & 3;
        const int r = qd * 32 + lane, ul = lane & 7;
        const int npr = priv >> 5, nsh = sh >> 5;                  // 32-column chunks of the private / shared zone
        const int nsh_h = nsh >> 1, npr_h = (npr - hh + 1) >> 1;   // this warp's share (chunks with index % 2 == hh)
        const int nj = nsh_h + npr_h;
        float* dp0 = reinterpret_cast<float*>(dpb) + (size_t)hh * 2 * DP_ONE;
        const int off4 = (ul & 1) * 16 + ((ul >> 1) & 1) * 8 + ((ul >> 2) & 1) * 4;   // columns this lane keeps after the u butterfly
        const int cbase = ((lane >> 3) & 1) * 16 + ((lane >> 4) & 1) * 8;              // ... after the t butterfly
        uint32_t q = 0, nchunk = 0;
        for (int run = blockIdx.x; run < nruns; run += gridDim.x) {
            const int bl = run / p.nTb, tb = run - bl * p.nTb, b = p.b0 + bl;
            const int Tn = p.xlen[b], Un = p.ylen[b] + 1;
            if (tb * BW_TT >= Tn) continue;
            const int nub = (Un + BW_UU - 1) / BW_UU;
            const int t = tb * BW_TT + (r >> 3);
            const float* erow = p.enc + ((size_t)b * p.maxT + min(t, p.maxT - 1)) * p.H;
            float accE[2][6][4];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int j = 0; j < 6; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) accE[a][j][i] = 0.f;
            for (int ub = 0; ub < nub; ++ub) {
                const int u = ub * BW_UU + ul;
                const float* prow = p.pred + ((size_t)b * p.maxU + min(u, p.maxU - 1)) * p.H;
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    if (pass >= NP) continue;
                    const uint32_t par = q & 1;
                    ptx::mbar_wait(&acc_full[par], (q >> 1) & 1);
                    ptx::tc_fence_after();
                    const int n0 = pass * NCZ;
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        if (j >= nj) continue;
                        const int c = j < nsh_h ? npr + 2 * j + hh : 2 * (j - nsh_h) + hh;      // chunk of this pass's NCZ columns
                        const uint32_t col = c < npr ? (par ? (uint32_t)p.odd_base : 0u) + 32u * c : 32u * c;
                        uint32_t v[32];
                        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(qd * 32) << 16) + col, v);
                        ptx::tmem_ld_wait();
                        if (j == nsh_h - 1 || j == nj - 1) {    // this warp is done with the shared zone / with the whole unit
                            ptx::tc_fence_before();
                            __syncwarp();
                            if (lane == 0) {
                                if (j == nsh_h - 1) ptx::mbar_arrive(shared_free);
                                if (j == nj - 1) ptx::mbar_arrive(&priv_free[par]);
                            }
                        }
                        const int h0 = n0 + 32 * c;
                        float g[32];
                        {
                            const float4* e4 = reinterpret_cast<const float4*>(erow + h0);
                            const float4* q4 = reinterpret_cast<const float4*>(prow + h0);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 e = __ldg(e4 + i), qq = __ldg(q4 + i);
                                const float z0 = ptx::tanh_approx(e.x + qq.x), z1 = ptx::tanh_approx(e.y + qq.y),
                                            z2 = ptx::tanh_approx(e.z + qq.z), z3 = ptx::tanh_approx(e.w + qq.w);
                                const float d0 = __uint_as_float(v[4 * i]), d1 = __uint_as_float(v[4 * i + 1]),
                                            d2 = __uint_as_float(v[4 * i + 2]), d3 = __uint_as_float(v[4 * i + 3]);
                                g[4 * i] = fmaf(-d0 * z0, z0, d0); g[4 * i + 1] = fmaf(-d1 * z1, z1, d1);
                                g[4 * i + 2] = fmaf(-d2 * z2, z2, d2); g[4 * i + 3] = fmaf(-d3 * z3, z3, d3);
                            }
                        }
                        // ---- sum over u (8 adjacent lanes): butterfly that halves the register set at every step
                        {
                            float a16[16], a8[8];
                            const bool h1 = lane & 1, h2 = lane & 2, h4 = lane & 4;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float send = h1 ? g[i] : g[16 + i], keep = h1 ? g[16 + i] : g[i];
                                a16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float send = h2 ? a16[i] : a16[8 + i], keep = h2 ? a16[8 + i] : a16[i];
                                a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                            }
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float send = h4 ? a8[i] : a8[4 + i], keep = h4 ? a8[4 + i] : a8[i];
                                accE[pass][j][i] += keep + __shfl_xor_sync(0xffffffffu, send, 4);
                            }
                        }
                        // ---- sum over t: 4 time steps inside the warp (lane bits 3, 4), then across the four quarter warps
                        {
                            float a16[16], a8[8];
                            const bool h8 = lane & 8, h16 = lane & 16;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float send = h8 ? g[i] : g[16 + i], keep = h8 ? g[16 + i] : g[i];
                                a16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float send = h16 ? a16[i] : a16[8 + i], keep = h16 ? a16[8 + i] : a16[i];
                                a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                            }
                            float* dp = dp0 + (size_t)(nchunk & 1) * DP_ONE;
                            float* w = dp + ((size_t)qslot * 8 + ul) * 36 + cbase;
                            *reinterpret_cast<float4*>(w) = make_float4(a8[0], a8[1], a8[2], a8[3]);
                            *reinterpret_cast<float4*>(w + 4) = make_float4(a8[4], a8[5], a8[6], a8[7]);
                            ptx::named_bar_sync(1 + hh, 128);
                            const int o = (qslot * 32 + lane) * 2, uo = o >> 5, co = o & 31;
                            float2 s = make_float2(0.f, 0.f);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 x = *reinterpret_cast<const float2*>(dp + ((size_t)k * 8 + uo) * 36 + co);
                                s.x += x.x; s.y += x.y;
                            }
                            const int uu = ub * BW_UU + uo;
                            if (uu < p.maxU)
                                *reinterpret_cast<float2*>(p.ppred + (((size_t)tb * p.nb + bl) * p.maxU + uu) * p.H + h0 + co) = s;
                            ++nchunk;
                        }
                    }
                    ++q;
                }
            }
            // ---- the run is complete: this lane owns (t, 4 columns per chunk) of d_enc
            if (t < p.maxT) {
                float* drow = p.d_enc + ((size_t)b * p.maxT + t) * p.H;
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    if (pass >= NP) continue;
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        if (j >= nj) continue;
                        const int c = j < nsh_h ? npr + 2 * j + hh : 2 * (j - nsh_h) + hh;
                        *reinterpret_cast<float4*>(drow + pass * NCZ + 32 * c + off4) =
                            make_float4(accE[pass][j][0], accE[pass][j][1], accE[pass][j][2], accE[pass][j][3]);
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

// d_pred[b,u,:] = sum over the t-blocks that intersect the utterance of its partial planes (0 for u >= U_b)
__global__ void __launch_bounds__(256) sum_pred_planes_kernel(const float4* __restrict__ ppred, const int* __restrict__ xlen,
                                                              const int* __restrict__ ylen, int b0, int nb, int maxU, int H4,
                                                              float4* __restrict__ d_pred) {
    const size_t n4 = (size_t)nb * maxU * H4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const int bl = (int)(i / ((size_t)maxU * H4)), u = (int)((i / H4) % maxU), b = b0 + bl;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u < ylen[b] + 1) {
            const int ntb = (xlen[b] + BW_TT - 1) / BW_TT;
#pragma unroll 4
            for (int k = 0; k < ntb; ++k) {
                const float4 v = ppred[(size_t)k * n4 + i];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        d_pred[(size_t)b0 * maxU * H4 + i] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// dW kernel
// ---------------------------------------------------------------------------------------------------------------
// CTA = (v-tile of DW_NV columns, h-item = one or two 128-row blocks of H, split of the K = lattice-row range).
// K step = half a tile (64 lattice rows: 8 time steps x 8 label positions).
struct DwWork { int vt, hb0, nblk, split, S; };
__device__ __forceinline__ DwWork dw_decode(const BwdParams& p, int cta) {
    const int per_vt = p.nD * p.S_d + p.nS * p.S_s;
    DwWork w;
    w.vt = cta / per_vt;
    int rem = cta - w.vt * per_vt;
    if (rem < p.nD * p.S_d) { const int it = rem / p.S_d; w.hb0 = 2 * it; w.nblk = 2; w.split = rem - it * p.S_d; w.S = p.S_d; }
    else { rem -= p.nD * p.S_d; w.hb0 = 2 * p.nD; w.nblk = 1; w.split = rem; w.S = p.S_s; }
    return w;
}

__global__ void __launch_bounds__(DW_THREADS, 1) bwd_dw_kernel(const __grid_constant__ CUtensorMap tmap_e, const BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t STAGE = 65536u;                     // A: 2 x [128 h x 64 k] K-major (32 KB) | B: 4 x [64 k x 64 v] MN-major (32 KB)
    float* tab = reinterpret_cast<float*>(smem + DW_STAGES * STAGE);        // 2 x { sc[64][8], specz[64], specw[64], labc[64] }
    constexpr int TAB_ONE = 64 * 8 + 3 * 64;
    float* dbs = tab + 2 * TAB_ONE;                                          // [4][256] column sums of the scaler row groups
    uint64_t* bars = reinterpret_cast<uint64_t*>(dbs + 4 * 256);
    uint64_t* b_full = bars;                       // [DW_STAGES] TMA -> scalers
    uint64_t* b_ready = b_full + DW_STAGES;        // scalers -> MMA
    uint64_t* a_ready = b_ready + DW_STAGES;       // producers -> MMA
    uint64_t* stage_empty = a_ready + DW_STAGES;   // MMA -> TMA, producers
    uint64_t* acc_full = stage_empty + DW_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const DwWork wk = dw_decode(p, blockIdx.x);
    const int v0 = wk.vt * DW_NV, Nv = min(DW_NV, p.V - v0), nbox = Nv >> 6, G = p.V >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < DW_STAGES; ++i) {
            ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_ready[i], 4);
            ptx::mbar_init(&a_ready[i], 4 * wk.nblk); ptx::mbar_init(&stage_empty[i], 1);
        }
        ptx::mbar_init(acc_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_ptr, TC_TMEM_COLS); ptx::tmem_relinquish(); }
    if (warp == 0 && lane == 0) ptx::prefetch_tmap(&tmap_e);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const long long cnt = *p.count;
    const int s_beg = (int)(cnt * wk.split / wk.S), s_end = (int)(cnt * (wk.split + 1) / wk.S);
    const int per_utt = p.nTb * p.nUb;

    if (warp == 0) {
        // ===================== TMA: B = kept numerators, boxes [64 rows x 64 v] (rows = K) =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int s = s_beg; s < s_end; ++s)
                for (int half = 0; half < 2; ++half) {
                    ptx::mbar_wait(&stage_empty[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx(&b_full[stage], (uint32_t)nbox * 8192u);
                    uint8_t* bs = smem + (size_t)stage * STAGE + 32768;
                    for (int j = 0; j < nbox; ++j)
                        ptx::tma_load_2d(bs + j * 8192, &tmap_e, &b_full[stage], v0 + 64 * j, s * 128 + half * 64);
                    if (++stage == DW_STAGES) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == 1) {
        // ===================== MMA: D[blk][128 h x Nv] += z^T[128 h x 64 k] . dl[64 k x Nv]  (A K-major, B MN-major) =====================
        const uint32_t idesc = ptx::umma_idesc_bf16(128, Nv, 0, 1);
        int stage = 0; uint32_t phase = 0, it = 0;
        for (int s = s_beg; s < s_end; ++s)
            for (int half = 0; half < 2; ++half, ++it) {
                ptx::mbar_wait(&b_ready[stage], phase);
                ptx::mbar_wait(&a_ready[stage], phase);
                ptx::tc_fence_after();
                const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * STAGE);
                const uint64_t bd = ptx::umma_desc_mn_sw128(sa + 32768u, 8192u);
                if (ptx::elect_one()) {
#pragma unroll
                    for (int blk = 0; blk < 2; ++blk) {
                        if (blk < wk.nblk) {
                            const uint64_t ad = ptx::umma_desc_k_sw128(sa + (uint32_t)blk * 16384u);
#pragma unroll
                            for (int k = 0; k < 4; ++k)   // K = 16 lattice rows per MMA: 32 B along A's rows, 16 x 128 B of B
                                ptx::umma_bf16(tmem_base + (uint32_t)blk * DW_NV, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 128),
                                               idesc, (uint32_t)((it | (uint32_t)k) != 0));
                        }
                    }
                    ptx::umma_commit(&stage_empty[stage]);
                    if (s == s_end - 1 && half == 1) ptx::umma_commit(acc_full);
                }
                __syncwarp();
                if (++stage == DW_STAGES) { stage = 0; phase ^= 1; }
            }
    } else if (warp < 6) {
        // ===================== B-stage scalers (warps 2-5) =====================
        constexpr float LOG2E = 1.4426950408889634f;
        const int r4 = threadIdx.x - 64, c = r4 & 31, rg = r4 >> 5;     // 16-byte chunk of the 512-byte row | group of 16 rows
        float colsum[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) colsum[i] = 0.f;
        // per-row inputs of the NEXT K step, fetched one step ahead by the threads that own a row (r4 < 64)
        float4 ncf = make_float4(-CUDART_INF_F, 0.f, 0.f, 0.f); int nlab = -1; float ngm[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ngm[i] = 0.f;
        auto fetch = [&](int s, int half) {
            if (r4 >= 64 || s >= s_end) return;
            const int tile = p.tile_of_slot ? p.tile_of_slot[s] : s;
            const int bl = tile / per_utt, rem = tile - bl * per_utt, b = p.b0 + bl;
            const int rr = half * 64 + r4;
            const int t = (rem / p.nUb) * BW_TT + (rr >> 3), u = (rem % p.nUb) * BW_UU + (rr & 7);
            const int Tn = p.xlen[b], Un = p.ylen[b] + 1;
            ncf = make_float4(-CUDART_INF_F, 0.f, 0.f, 0.f); nlab = -1;
            if (t < Tn && u < Un) {
                ncf = p.coef[((long long)b * p.maxT + t) * p.maxU + u];
                if (u < Un - 1) nlab = p.labels[(size_t)b * (p.maxU - 1) + u];
            }
            const float* gmr = p.gm + ((size_t)s * G + (v0 >> 5)) * 128 + rr;
#pragma unroll
            for (int i = 0; i < 8; ++i) ngm[i] = (i < 2 * nbox) ? gmr[(size_t)i * 128] : 0.f;
        };
        fetch(s_beg, 0);
        int stage = 0; uint32_t phase = 0, it = 0;
        for (int s = s_beg; s < s_end; ++s)
            for (int half = 0; half < 2; ++half, ++it) {
                float* tb = tab + (size_t)(it & 1) * TAB_ONE;
                if (r4 < 64) {
                    const float kd2 = ncf.x * LOG2E;
#pragma unroll
                    for (int i = 0; i < 8; ++i) tb[r4 * 8 + i] = ncf.y * ptx::ex2_approx(ngm[i] + kd2);   // (invalid row: g = 0, kd = -inf -> 0)
                    tb[512 + r4] = ncf.z; tb[576 + r4] = ncf.w;
                    reinterpret_cast<int*>(tb)[640 + r4] = nlab >= 0 ? nlab - v0 : -100000;
                }
                if (half == 0) fetch(s, 1); else fetch(s + 1, 0);
                ptx::named_bar_sync(3, 128);
                ptx::mbar_wait(&b_full[stage], phase);
                uint8_t* bs = smem + (size_t)stage * STAGE + 32768;
                if ((c >> 3) < nbox) {
                    const int dbk = p.blank - v0 - c * 8;
#pragma unroll 4
                    for (int rr = 0; rr < 16; ++rr) {
                        const int k = rg * 16 + rr;
                        uint4* ptr = reinterpret_cast<uint4*>(bs + (c >> 3) * 8192 + k * 128 + (((c & 7) ^ (k & 7)) << 4));
                        const uint4 x = *ptr;
                        const int dl = reinterpret_cast<const int*>(tb)[640 + k] - c * 8;
                        *ptr = scale_chunk(x, tb[k * 8 + (c >> 2)], dbk, tb[512 + k], dl, tb[576 + k], colsum);
                    }
                }
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&b_ready[stage]);
                if (++stage == DW_STAGES) { stage = 0; phase ^= 1; }
            }
        // db: column sums of this CTA's K range (only the first h-item of a v-tile reports them)
        if (wk.hb0 == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dbs[rg * 256 + c * 8 + i] = colsum[i];
            ptx::named_bar_sync(3, 128);
            for (int v = r4; v < Nv; v += 128) {
                const float sum = dbs[v] + dbs[256 + v] + dbs[512 + v] + dbs[768 + v];
                float* dst = p.dbp + (size_t)wk.split * p.V + v0 + v;
                *dst = p.accumulate ? *dst + sum : sum;
            }
        }
    } else {
        // ===================== z producers (warps 6-13): thread = column h of the joint; then the epilogue =====================
        const int ptid = threadIdx.x - 192, blk = ptid >> 7, hl = ptid & 127;
        const int h = (wk.hb0 + blk) * 128 + hl;
        const bool active = blk < wk.nblk, hv = h < p.H;
        if (active) {
            int stage = 0; uint32_t phase = 0;
            for (int s = s_beg; s < s_end; ++s) {
                const int tile = p.tile_of_slot ? p.tile_of_slot[s] : s;
                const int bl = tile / per_utt, rem = tile - bl * per_utt, b = p.b0 + bl;
                const int t0 = (rem / p.nUb) * BW_TT, u0 = (rem % p.nUb) * BW_UU;
                const int Tn = p.xlen[b], Un = p.ylen[b] + 1;
                float pv[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) pv[k] = hv ? __ldg(p.pred + ((size_t)b * p.maxU + min(u0 + k, p.maxU - 1)) * p.H + h) : 0.f;
                for (int half = 0; half < 2; ++half) {
                    float ev[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) ev[k] = hv ? __ldg(p.enc + ((size_t)b * p.maxT + min(t0 + half * 8 + k, p.maxT - 1)) * p.H + h) : 0.f;
                    ptx::mbar_wait(&stage_empty[stage], phase ^ 1);
                    uint8_t* arow = smem + (size_t)stage * STAGE + blk * 16384 + hl * 128;
#pragma unroll
                    for (int tl = 0; tl < 8; ++tl) {
                        const bool tv = hv && (t0 + half * 8 + tl) < Tn;
                        float z[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) z[k] = (tv && (u0 + k) < Un) ? ptx::tanh_approx(ev[tl] + pv[k]) : 0.f;
                        *reinterpret_cast<uint4*>(arow + ((tl ^ (hl & 7)) << 4)) =
                            make_uint4(ptx::pack_bf16x2(z[0], z[1]), ptx::pack_bf16x2(z[2], z[3]), ptx::pack_bf16x2(z[4], z[5]),
                                       ptx::pack_bf16x2(z[6], z[7]));
                    }
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&a_ready[stage]);
                    if (++stage == DW_STAGES) { stage = 0; phase ^= 1; }
                }
            }
            // ---- epilogue: this warp's 32 rows (h) of block `blk`, Nv columns -> the split's partial plane
            if (s_end > s_beg) {
                ptx::mbar_wait(acc_full, 0);
                ptx::tc_fence_after();
                const int qd = warp & 3, hr = (wk.hb0 + blk) * 128 + qd * 32 + lane;
                float* dst = p.dWp + ((size_t)wk.split * p.Hrows + hr) * p.V + v0;
                for (int j = 0; j < (Nv >> 5); ++j) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(blk * DW_NV + j * 32), v);
                    ptx::tmem_ld_wait();
                    if (hr < p.Hrows) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float4 o = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                   __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                            float4* d4 = reinterpret_cast<float4*>(dst + j * 32 + i * 4);
                            if (p.accumulate) { const float4 x = *d4; o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w; }
                            *d4 = o;
                        }
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

}  // namespace rb
