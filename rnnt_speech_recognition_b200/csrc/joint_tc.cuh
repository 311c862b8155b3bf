// joint_tc.cuh -- bf16 tensor-core (tcgen05 / TMEM / TMA) path of the fused joint + loss: shared definitions
// (tile geometry, kernel parameters, tensor maps, scratch layout), the FIRST-generation kernel, the backward's
// reduction kernels and the host-side orchestration (tc_forward / tc_backward / tc_dispatch).
// The default kernel is joint_tc3_kernel (joint_tc3.cuh); joint_tc_kernel below is generation 1
// (RNNTB200_TC_VARIANT=1), also the fallback when z does not fit tensor memory (H > 768).
//
// joint_tc_kernel<MODE>: one persistent CTA per SM, each looping over 128-cell lattice tiles
// (TT time steps x UU label positions of one utterance).  Per tile:
//   * 8 producer warps form the A operand  z = tanh(enc[b,t,:] + pred[b,u,:])  (bf16) straight into
//     shared memory in the canonical K-major SWIZZLE_128B layout (H/64 K-blocks of 128x64) -- the
//     (B,T,U,H) activation tensor of model.py:158-163 never exists in HBM;
//   * one TMA thread streams W^T (bf16, V x H, K-major) tiles [NC x 64] through an mbarrier ring;
//   * one MMA thread issues tcgen05.mma (M=128, N=NC, K=16) into double-buffered TMEM accumulators;
//   * 4 epilogue warps read the accumulators with tcgen05.ld (thread = lattice cell) and
//       MODE 0 (forward):  add bias, ONLINE log-sum-exp across the V chunks, pick logit[blank] and
//                          logit[label_u]  ->  lse (natural order), lp_blank / lp_label (skewed planes)
//                          -- the (B,T,U,V) logits of model.py:165-166 never reach HBM;
//       MODE 1 (backward): dlogit = g*exp(x + kd) for all columns, then the blank / label entries are overwritten
//                          with the values cell_coef_kernel precomputed -> bf16 rows of the dlogits workspace
//                          (+ the bf16 z rows the dW GEMM consumes).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <mutex>
#include <unordered_map>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>

#ifndef RNNTB200_DEFAULT_TC_VARIANT
#define RNNTB200_DEFAULT_TC_VARIANT 3
#endif

#include "../../include/rnnt_b200.h"
#include "kernels_simt.cuh"
#include "ptx.cuh"
#include "timing.cuh"

namespace rb {

constexpr int TC_THREADS = 320;      // warps 0-3 epilogue+producer, 4-7 producer, 8 TMA, 9 MMA
constexpr int TC_MAX_KB = 16;        // H <= 1024
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_TMEM_COLS = 512;

struct TcGeom {
    int TT, UU, nTb, nUb;  // tile = TT x UU cells (TT*UU == 128); tiles per utterance
    int NC, NCH, KB, stages;
    size_t smem_bytes;
    bool ok;
};

inline TcGeom tc_geometry(int maxT, int maxU, int H, int V) {
    TcGeom g{};
    g.ok = false;
    if (H % 64 || V % 64 || H > 64 * TC_MAX_KB) return g;
    // tile shape: least padded lattice cells first; among equals the shape that loads the fewest enc + pred rows per
    // tile (TT + UU): every tile re-reads its TT enc rows and UU pred rows (fp32, H wide) from L2, and the L2 -> SM
    // stream is what bounds the fused kernel (1 x 128 tiles: 330 KB per tile; 8 x 16: 61 KB).  TT <= 16 (enc TMA box).
    int best = 128;
    long long best_pad = 1ll << 60;
    int best_rows = 1 << 30;
    for (int uu = 128; uu >= 8; uu >>= 1) {
        const int tt = 128 / uu;
        const long long pad = (long long)((maxU + uu - 1) / uu * uu) * ((maxT + tt - 1) / tt * tt);
        // (ties -> the narrower tile: 16 x 8 is the shape the single-pass reduction is written for)
        if (pad < best_pad || (pad == best_pad && tt + uu <= best_rows)) { best_pad = pad; best_rows = tt + uu; best = uu; }
    }
    {   // RNNTB200_TILE_UU = 8..128 forces the tile width (A/B measurements)
        static int force = -1;
        if (force < 0) { const char* e = getenv("RNNTB200_TILE_UU"); force = e ? atoi(e) : 0; }
        if (force >= 8 && force <= 128 && (force & (force - 1)) == 0) best = force;
    }
    g.UU = best;
    g.TT = 128 / best;
    g.nTb = (maxT + g.TT - 1) / g.TT;
    g.nUb = (maxU + g.UU - 1) / g.UU;
    g.KB = H / 64;
    const size_t zbytes = (size_t)g.KB * 16384, limit = 232448 - 1024 /*alignment slack*/ - 512 /*barriers*/;
    for (int nc = 256; nc >= 64; nc >>= 1) {
        if (V % nc) continue;
        for (int st = TC_MAX_STAGES; st >= 2; --st) {
            if (zbytes + (size_t)st * nc * 128 <= limit) {
                g.NC = nc; g.NCH = V / nc; g.stages = st;
                g.smem_bytes = zbytes + (size_t)st * nc * 128 + 512 + 1024;
                g.ok = true;
                return g;
            }
        }
    }
    return g;
}

struct JointTcParams {
    const float* enc; const float* pred; const float* bias;
    const int* labels; const int* xlen; const int* ylen;
    int B, maxT, maxU, H, V, blank;
    int TT, UU, nTb, nUb, NC, NCH, KB, stages;
    long long SK;
    int b0, nb;                 // utterance range of this launch
    const int* slot;            // optional tile -> compact row-block map of this launch (valid tiles only); NULL = tile order
    int zld;                    // row stride (elements) of zb: tc_zld(H) = H + 16 (column H carries the ones column for db; 32-byte aligned rows)
    int nbuf, swap, ks, dbg;    // v2 kernel: TMEM accumulator buffers; bf16-pair order of TMEM A; K-blocks per W stage; bring-up switches
    float* lse; float* lpb; float* lpl;              // MODE 0 outputs
    const float4* coef; __nv_bfloat16* dl; __nv_bfloat16* zb;  // MODE 1: coefficients in, dlogits / z rows out
    // MODE 2 (forward that KEEPS its activations): dl receives the softmax numerators 2^(y - gm) as fp16 (same rows,
    // same bytes as the bf16 dlogits that dl_from_kept_kernel later writes over them), gm the running maximum each
    // 32-column group was taken against, zb the tanh outputs.
    float* gm;
};

// zb row pitch: H + 16 elements keeps every row 32-byte aligned (256-bit stores); the dW GEMM reads H + 8 columns of it
inline int tc_zld(int H) { return H + 16; }
struct TileInfo { int b, t0, u0, Tn, Un; bool valid; };
__device__ __forceinline__ TileInfo decode_tile(const JointTcParams& p, int tile) {
    TileInfo ti;
    const int per_utt = p.nTb * p.nUb;
    const int bl = tile / per_utt, rem = tile - bl * per_utt;
    ti.b = p.b0 + bl;
    ti.t0 = (rem / p.nUb) * p.TT;
    ti.u0 = (rem % p.nUb) * p.UU;
    ti.Tn = p.xlen[ti.b];
    ti.Un = p.ylen[ti.b] + 1;
    ti.valid = ti.t0 < ti.Tn && ti.u0 < ti.Un;
    return ti;
}

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) joint_tc_kernel(const __grid_constant__ CUtensorMap tmap_wt,
                                                                 const JointTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int KB = p.KB, NC = p.NC, NCH = p.NCH, stages = p.stages;
    uint8_t* zs = smem;                                   // KB x [128 x 64] bf16, SW128 K-major
    uint8_t* wsm = smem + (size_t)KB * 16384;             // stages x [NC x 64] bf16, SW128 K-major (TMA)
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + (size_t)stages * NC * 128);
    uint64_t* z_full = bars;                              // [TC_MAX_KB]  producers -> MMA
    uint64_t* z_free = bars + TC_MAX_KB;                  //              MMA -> producers (tile's MMAs retired)
    uint64_t* w_full = z_free + 1;                        // [stages]     TMA -> MMA
    uint64_t* w_empty = w_full + TC_MAX_STAGES;           // [stages]     MMA -> TMA
    uint64_t* acc_full = w_empty + TC_MAX_STAGES;         // [2]          MMA -> epilogue
    uint64_t* acc_empty = acc_full + 2;                   // [2]          epilogue -> MMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_MAX_KB; ++i) ptx::mbar_init(&z_full[i], 8);
        ptx::mbar_init(z_free, 1);
        for (int i = 0; i < TC_MAX_STAGES; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 4); }
        ptx::fence_barrier_init();
    }
    if (warp == 9) { ptx::tmem_alloc(tmem_ptr, TC_TMEM_COLS); ptx::tmem_relinquish(); }
    if (warp == 8 && lane == 0) ptx::prefetch_tmap(&tmap_wt);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int ntiles = p.nb * p.nTb * p.nUb;

    if (warp == 8) {
        // ===================== TMA producer: W^T tiles [NC rows (v) x 64 (k)] =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                if (!decode_tile(p, tile).valid) continue;
                for (int c = 0; c < NCH; ++c)
                    for (int kb = 0; kb < KB; ++kb) {
                        ptx::mbar_wait(&w_empty[stage], phase ^ 1);
                        ptx::mbar_arrive_expect_tx(&w_full[stage], (uint32_t)NC * 128);
                        ptx::tma_load_2d(wsm + (size_t)stage * NC * 128, &tmap_wt, &w_full[stage], kb * 64, c * NC);
                        if (++stage == stages) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer =====================
        // whole warp convergent, one elected lane issues (operands stay in uniform registers)
        const uint32_t idesc = ptx::umma_idesc_bf16(128, NC);
        int stage = 0; uint32_t phase = 0, g = 0, it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            if (!decode_tile(p, tile).valid) continue;
            for (int c = 0; c < NCH; ++c, ++g) {
                const uint32_t buf = g & 1;
                ptx::mbar_wait(&acc_empty[buf], ((g >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * NC;
                for (int kb = 0; kb < KB; ++kb) {
                    if (c == 0) ptx::mbar_wait(&z_full[kb], it & 1);
                    ptx::mbar_wait(&w_full[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t ad = ptx::umma_desc_k_sw128(ptx::smem_u32(zs + (size_t)kb * 16384));
                    const uint64_t bd = ptx::umma_desc_k_sw128(ptx::smem_u32(wsm + (size_t)stage * NC * 128));
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            ptx::umma_bf16(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc,
                                           (uint32_t)((kb | k) != 0));
                        ptx::umma_commit(&w_empty[stage]);
                        if (kb == KB - 1) ptx::umma_commit(&acc_full[buf]);
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
            if (ptx::elect_one()) ptx::umma_commit(z_free);
            __syncwarp();
            ++it;
        }
    } else {
        // ===================== 8 compute warps: z producers (all) + epilogue (warps 0-3) =====================
        constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
        uint32_t g = 0, it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const TileInfo ti = decode_tile(p, tile);
            if (!ti.valid) {
                if (MODE == 1 && !p.slot) {  // uncompacted rows: the plain GEMMs reduce over ALL rows, padding tiles must read as zero
                    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
                    uint4* d4 = reinterpret_cast<uint4*>(p.dl + (size_t)tile * 128 * p.V);
                    for (int i = threadIdx.x; i < 128 * p.V / 8; i += 256) d4[i] = z4;
                    if (p.zb) {
                        const int h8 = p.H / 8;
                        for (int i = threadIdx.x; i < 128 * h8; i += 256)
                            *reinterpret_cast<uint4*>(p.zb + ((size_t)tile * 128 + i / h8) * p.zld + (i % h8) * 8) = z4;
                    }
                }
                continue;
            }
            const size_t rowbase = (size_t)(p.slot ? p.slot[tile] : tile) * 128;   // row block of this tile in dl / zb
            // ---- A operand: z = tanh(enc + pred) -> bf16, SW128 K-major, one 16-byte chunk per thread-task
            // Per-thread task geometry is fixed for the tile: 4 row-passes x one 16-byte chunk column.
            uint32_t eo[4], qo[4], soff[4];   // float4-unit offsets into enc / pred, byte offset into a K block
            bool ok[4];
            const float4* enc4 = reinterpret_cast<const float4*>(p.enc);
            const float4* pred4 = reinterpret_cast<const float4*>(p.pred);
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
                const int r = pass * 32 + warp * 4 + (lane >> 3), ch = lane & 7;
                const int t = ti.t0 + r / p.UU, u = ti.u0 + r % p.UU;
                ok[pass] = t < ti.Tn && u < ti.Un;
                eo[pass] = (uint32_t)((((size_t)ti.b * p.maxT + (ok[pass] ? t : 0)) * p.H + ch * 8) >> 2);
                qo[pass] = (uint32_t)((((size_t)ti.b * p.maxU + (ok[pass] ? u : 0)) * p.H + ch * 8) >> 2);
                soff[pass] = r * 128 + ((ch ^ (r & 7)) << 4);
            }
            // Global loads of K-block kb+1 are issued before the tanh work of K-block kb (two register
            // buffers, loop unrolled by two) so that their L2 latency overlaps the MUFU work.
            float4 bufA[16], bufB[16];
            auto issue = [&](int kb, float4* buf) {
#pragma unroll
                for (int pass = 0; pass < 4; ++pass) {
                    const float4* e = enc4 + eo[pass] + kb * 16;   // 64 floats per K block = 16 float4
                    const float4* q = pred4 + qo[pass] + kb * 16;
                    buf[pass * 4 + 0] = __ldg(e); buf[pass * 4 + 1] = __ldg(e + 1);
                    buf[pass * 4 + 2] = __ldg(q); buf[pass * 4 + 3] = __ldg(q + 1);
                }
            };
            auto produce = [&](int kb, const float4* buf) {
#pragma unroll
                for (int pass = 0; pass < 4; ++pass) {
                    const float4 e0 = buf[pass * 4], e1 = buf[pass * 4 + 1], q0 = buf[pass * 4 + 2], q1 = buf[pass * 4 + 3];
                    uint4 packed = make_uint4(0u, 0u, 0u, 0u);
                    if (ok[pass]) {
                        packed.x = ptx::pack_bf16x2(ptx::tanh_approx(e0.x + q0.x), ptx::tanh_approx(e0.y + q0.y));
                        packed.y = ptx::pack_bf16x2(ptx::tanh_approx(e0.z + q0.z), ptx::tanh_approx(e0.w + q0.w));
                        packed.z = ptx::pack_bf16x2(ptx::tanh_approx(e1.x + q1.x), ptx::tanh_approx(e1.y + q1.y));
                        packed.w = ptx::pack_bf16x2(ptx::tanh_approx(e1.z + q1.z), ptx::tanh_approx(e1.w + q1.w));
                    }
                    *reinterpret_cast<uint4*>(zs + (size_t)kb * 16384 + soff[pass]) = packed;
                    if (MODE == 1 && p.zb) {
                        const int r = pass * 32 + warp * 4 + (lane >> 3);
                        *reinterpret_cast<uint4*>(p.zb + (rowbase + r) * p.zld + kb * 64 + (lane & 7) * 8) = packed;
                    }
                }
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&z_full[kb]);
            };
            issue(0, bufA);
            ptx::mbar_wait(z_free, (it & 1) ^ 1);
            for (int kb = 0; kb < KB; kb += 2) {
                if (kb + 1 < KB) issue(kb + 1, bufB);
                produce(kb, bufA);
                if (kb + 1 < KB) {
                    if (kb + 2 < KB) issue(kb + 2, bufA);
                    produce(kb + 1, bufB);
                }
            }
            // ---- epilogue: thread = lattice cell (TMEM lane), warps 0-3 cover lanes 0..127
            if (warp < 4) {
                const int r = warp * 32 + lane;
                const int t = ti.t0 + r / p.UU, u = ti.u0 + r % p.UU;
                const bool rv = t < ti.Tn && u < ti.Un;
                const int lab = (rv && u < ti.Un - 1) ? p.labels[(size_t)ti.b * (p.maxU - 1) + u] : -1;
                const long long cell = ((long long)ti.b * p.maxT + t) * p.maxU + u;
                float m2 = -CUDART_INF_F, s = 0.f, yb = 0.f, yl = 0.f;  // MODE 0 state (log2 domain)
                float kd2 = -CUDART_INF_F, cg = 0.f, csb = 0.f, csl = 0.f;  // MODE 1: exponent offset, scale, final dl[blank], final dl[label]
                if (MODE == 1 && rv) {
                    const float4 cf = p.coef[cell];
                    kd2 = cf.x * LOG2E; cg = cf.y; csb = cf.z; csl = cf.w;
                }
                const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
                for (int c = 0; c < NCH; ++c, ++g) {
                    const uint32_t buf = g & 1;
                    ptx::mbar_wait(&acc_full[buf], (g >> 1) & 1);
                    ptx::tc_fence_after();
                    for (int j = 0; j < NC / 32; ++j) {
                        uint32_t v[32];
                        ptx::tmem_ld_32x32(lane_addr + buf * NC + j * 32, v);
                        ptx::tmem_ld_wait();
                        const int col0 = c * NC + j * 32;
                        // one coalesced bias load per warp, broadcast lane-by-lane (a per-element LDG would
                        // saturate the LSU long before the MMA pipe)
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
                            if (p.blank >= col0 && p.blank < col0 + 32) {  // uniform
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
            } else {
                g += NCH;
            }
            ++it;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 9) ptx::tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

// W (H,V) fp32 -> Wt (V,H) bf16 [B operand of the logits GEMM, K=h] and Wb (H,V) bf16 [B operand of dZ, K=v]
__global__ void __launch_bounds__(256) convert_w_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ Wt,
                                                        __nv_bfloat16* __restrict__ Wb, int H, int V) {
    __shared__ float tile[32][33];
    const int v0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int h = h0 + i, v = v0 + tx;
        const float w = (h < H && v < V) ? W[(size_t)h * V + v] : 0.f;
        tile[i][tx] = w;
        if (h < H && v < V) Wb[(size_t)h * V + v] = __float2bfloat16(w);
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int v = v0 + i, h = h0 + tx;
        if (h < H && v < V) Wt[(size_t)v * H + h] = __float2bfloat16(tile[tx][i]);
    }
}

// ------------------------------------------------------------------------------------------
// Reductions of the backward (tile-ordered rows: row = tile*128 + (t%TT)*UU + (u%UU)).
//   g = dZ * sech^2(enc+pred), with sech^2 evaluated from exp (relative accuracy near |z| -> 1,
//   where 1 - tanh^2 computed from a rounded tanh would lose all its digits)
// ------------------------------------------------------------------------------------------
struct RowMap { int TT, UU, nTb, nUb, b0, lgUU; const int* slot; };   // TT*UU == 128, both powers of two
__device__ __forceinline__ size_t tile_row(const RowMap& m, int b, int t, int u) {
    const int lgTT = 7 - m.lgUU;
    size_t q = ((size_t)(b - m.b0) * m.nTb + (t >> lgTT)) * m.nUb + (u >> m.lgUU);
    if (m.slot) q = (size_t)m.slot[q];     // compacted: only valid tiles own rows (callers only ask for valid cells)
    return q * 128 + ((t & (m.TT - 1)) << m.lgUU) + (u & (m.UU - 1));
}
// slot[tile] = rank of the tile among the VALID tiles of this launch (-1 if it lies in the padding); *count = #valid.
// One block; a serial-over-chunks block scan is plenty for the <= ~1e5 tiles of a launch.
__global__ void __launch_bounds__(1024) tile_compact_kernel(const int* __restrict__ xlen, const int* __restrict__ ylen,
                                                            int b0, int ntiles, int nTb, int nUb, int TT, int UU,
                                                            int* __restrict__ slot, int* __restrict__ count) {
    __shared__ int warp_sums[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int per_utt = nTb * nUb, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t0 = 0; t0 < ntiles; t0 += 1024) {
        const int tile = t0 + threadIdx.x;
        int v = 0;
        if (tile < ntiles) {
            const int bl = tile / per_utt, rem = tile - bl * per_utt, b = b0 + bl;
            v = ((rem / nUb) * TT < xlen[b] && (rem % nUb) * UU < ylen[b] + 1) ? 1 : 0;
        }
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += n; }
            warp_sums[lane] = wi - w;   // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int excl = base + warp_sums[warp] + incl - v;
        if (tile < ntiles) slot[tile] = v ? excl : -1;
        __syncthreads();
        if (threadIdx.x == 1023) base = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}
inline int ilog2(int x) { int l = 0; while ((1 << l) < x) ++l; return l; }
// g = dZ * sech^2(enc+pred) reduced over u (d_enc) and over t (d_pred).  Both kernels stream WHOLE rows of the
// bf16 dZ (H elements, contiguous): block = H/8 threads x 16-byte loads, one row per loop iteration, unrolled for
// memory-level parallelism.
struct F8 { float v[8]; };
__device__ __forceinline__ F8 load_bf16x8(const __nv_bfloat16* p) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    F8 o;
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { o.v[2 * i] = __uint_as_float(w[i] << 16); o.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
    return o;
}
__device__ __forceinline__ F8 load_f32x8(const float* p) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    F8 o; o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w; o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
    return o;
}
__device__ __forceinline__ void store_f32x8(float* p, const F8& a) {
    *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
// Pass 1, grid (maxT, nb): g = dZ * sech^2(enc+pred) is computed ONCE, written back in place (bf16) and summed over u:
//   d_enc[b,t,:] = sum_u g
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* p, const F8& a) {
    *reinterpret_cast<uint4*>(p) = make_uint4(ptx::pack_bf16x2(a.v[0], a.v[1]), ptx::pack_bf16x2(a.v[2], a.v[3]),
                                              ptx::pack_bf16x2(a.v[4], a.v[5]), ptx::pack_bf16x2(a.v[6], a.v[7]));
}
__global__ void __launch_bounds__(128) denc_rows_kernel(__nv_bfloat16* __restrict__ dz, const float* __restrict__ enc,
                                                        const float* __restrict__ pred, const int* __restrict__ xlen,
                                                        const int* __restrict__ ylen, RowMap m, int maxT, int maxU,
                                                        int H, float* __restrict__ d_enc) {
    const int t = blockIdx.x, b = m.b0 + blockIdx.y, h = threadIdx.x * 8;
    if (h >= H) return;
    const int Tn = xlen[b], Un = ylen[b] + 1;
    F8 acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
    if (t < Tn) {
        const F8 e = load_f32x8(enc + ((size_t)b * maxT + t) * H + h);
        const float* prow = pred + (size_t)b * maxU * H + h;
        // rows of one u-block of the tile grid are contiguous: one tile_row() per block, pointer increments inside.
        // sech^2 = 1 - tanh^2 with the same tanh.approx the forward used for z: one MUFU op and two FMAs per element
        // (this kernel is issue-bound, not bandwidth-bound).
        for (int u0 = 0; u0 < Un; u0 += m.UU) {
            __nv_bfloat16* row = dz + tile_row(m, b, t, u0) * H + h;
            const int un = min(m.UU, Un - u0);
#pragma unroll 4
            for (int ul = 0; ul < un; ++ul, row += H, prow += H) {
                const F8 d = load_bf16x8(row);
                const F8 q = load_f32x8(prow);
                F8 g;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float z = ptx::tanh_approx(e.v[i] + q.v[i]);
                    g.v[i] = fmaf(-d.v[i] * z, z, d.v[i]);
                    acc.v[i] += g.v[i];
                }
                store_bf16x8(row, g);
            }
        }
    }
    store_f32x8(d_enc + ((size_t)b * maxT + t) * H + h, acc);
}
// Pass 2, grid (maxU, nb): d_pred[b,u,:] = sum_t g   (pure sum over the rows pass 1 rewrote)
__global__ void __launch_bounds__(128) dpred_rows_kernel(const __nv_bfloat16* __restrict__ g, const int* __restrict__ xlen,
                                                         const int* __restrict__ ylen, RowMap m, int maxU, int H,
                                                         float* __restrict__ d_pred) {
    const int u = blockIdx.x, b = m.b0 + blockIdx.y, h = threadIdx.x * 8;
    if (h >= H) return;
    const int Tn = xlen[b], Un = ylen[b] + 1;
    F8 acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
    if (u < Un) {
#pragma unroll 8
        for (int t = 0; t < Tn; ++t) {
            const F8 d = load_bf16x8(g + tile_row(m, b, t, u) * H + h);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc.v[i] += d.v[i];
        }
    }
    store_f32x8(d_pred + ((size_t)b * maxU + u) * H + h, acc);
}
// Single-pass reduction for 16 x 8 tiles: g = dZ * (1 - tanh^2) is formed once, in registers, and never written back.
// A block owns one u-block column (8 label positions) of one utterance and sweeps a segment of its tiles along t, so
// it reads WHOLE tiles (128 contiguous rows) one after the other; thread <-> EIGHT columns of H (16-byte loads: the
// number of outstanding load requests per SM is what limits a streaming kernel, so each request should carry 512
// bytes per warp -- the same design with 4-byte loads ran at 1 TB/s).  Per time step the 8 rows are summed over u
// straight into a partial plane of d_enc (one plane per u-block); the 8 column sums over t accumulate in registers
// across the sweep and go to a partial plane of d_pred (one per segment); sum_planes_kernel adds the planes.
// HBM traffic: one read of the bf16 dZ plus the small fp32 planes, instead of read + write-back + second read.
constexpr int RED_THREADS = 96;                    // H <= 640 on the tensor-core path: 80 threads x 8 columns
constexpr int RED_MIN_BLOCKS = 2368;               // ~4 waves of 4 blocks per SM
constexpr int RED_UU = 8, RED_TT = 16;
__global__ void __launch_bounds__(RED_THREADS, 4) dencpred_tiles_kernel(
    const __nv_bfloat16* __restrict__ dz, const float* __restrict__ enc, const float* __restrict__ pred,
    const int* __restrict__ xlen, const int* __restrict__ ylen, RowMap m, int maxT, int maxU, int H, int tiles_per_seg,
    float* __restrict__ penc, float* __restrict__ ppred) {
    __shared__ float4 q_s[RED_UU * 2 * RED_THREADS];   // the block's pred rows, [k][half][thread]: thread-private columns
    const int ub = blockIdx.x, seg = blockIdx.y, bl = blockIdx.z, nb = gridDim.z, b = m.b0 + bl;
    const int h = threadIdx.x * 8, u0 = ub * RED_UU;
    if (h >= H) return;                             // (no block-wide barrier below)
    const int Tn = xlen[b], Un = ylen[b] + 1;
    float4* qs = q_s + threadIdx.x;
#pragma unroll
    for (int k = 0; k < RED_UU; ++k) {
        const float* src = pred + ((size_t)b * maxU + min(u0 + k, maxU - 1)) * H + h;   // (rows with u >= U_b carry dZ == 0)
        qs[(2 * k) * RED_THREADS] = *reinterpret_cast<const float4*>(src);
        qs[(2 * k + 1) * RED_THREADS] = *reinterpret_cast<const float4*>(src + 4);
    }
    float accP[RED_UU][8];
#pragma unroll
    for (int k = 0; k < RED_UU; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) accP[k][i] = 0.f;
    const bool ub_valid = u0 < Un;
    const int tb_end = min((seg + 1) * tiles_per_seg, m.nTb);
    for (int tb = seg * tiles_per_seg; tb < tb_end; ++tb) {
        const int t0 = tb * RED_TT;
        float* pe = penc + (((size_t)ub * nb + bl) * maxT + t0) * H + h;
        if (!(ub_valid && t0 < Tn)) {               // tile outside the valid lattice: its d_enc share is zero
            for (int j = 0; j < RED_TT; ++j)
                if (t0 + j < maxT) {
                    *reinterpret_cast<float4*>(pe + (size_t)j * H) = make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(pe + (size_t)j * H + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            continue;
        }
        size_t q = ((size_t)bl * m.nTb + tb) * m.nUb + ub;
        if (m.slot) q = (size_t)m.slot[q];
        const __nv_bfloat16* base = dz + q * 128 * H + h;
#pragma unroll 2
        for (int j = 0; j < RED_TT; ++j) {
            uint4 d[RED_UU];
#pragma unroll
            for (int k = 0; k < RED_UU; ++k)        // the 8 rows of this time step: 8 independent 16-byte requests
                d[k] = *reinterpret_cast<const uint4*>(base + (size_t)(j * RED_UU + k) * H);
            const int t = min(t0 + j, maxT - 1);    // (rows with t >= T_b carry dZ == 0)
            const float* ep = enc + ((size_t)b * maxT + t) * H + h;
            const float4 e0 = __ldg(reinterpret_cast<const float4*>(ep)), e1 = __ldg(reinterpret_cast<const float4*>(ep + 4));
            const float e[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
            float aE[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) aE[i] = 0.f;
#pragma unroll
            for (int k = 0; k < RED_UU; ++k) {
                const float4 q0 = qs[(2 * k) * RED_THREADS], q1 = qs[(2 * k + 1) * RED_THREADS];
                const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                const uint32_t w[4] = {d[k].x, d[k].y, d[k].z, d[k].w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float dv = (i & 1) ? __uint_as_float(w[i >> 1] & 0xffff0000u) : __uint_as_float(w[i >> 1] << 16);
                    const float z = ptx::tanh_approx(e[i] + qv[i]);
                    const float g = fmaf(-dv * z, z, dv);
                    aE[i] += g;
                    accP[k][i] += g;
                }
            }
            if (t0 + j < maxT) {
                *reinterpret_cast<float4*>(pe + (size_t)j * H) = make_float4(aE[0], aE[1], aE[2], aE[3]);
                *reinterpret_cast<float4*>(pe + (size_t)j * H + 4) = make_float4(aE[4], aE[5], aE[6], aE[7]);
            }
        }
    }
    float* pp = ppred + (((size_t)seg * nb + bl) * maxU + u0) * H + h;
#pragma unroll
    for (int k = 0; k < RED_UU; ++k)
        if (u0 + k < maxU) {
            *reinterpret_cast<float4*>(pp + (size_t)k * H) = make_float4(accP[k][0], accP[k][1], accP[k][2], accP[k][3]);
            *reinterpret_cast<float4*>(pp + (size_t)k * H + 4) = make_float4(accP[k][4], accP[k][5], accP[k][6], accP[k][7]);
        }
}
// out[i] = sum_k part[k][i] over `nplanes` planes of `n4` float4 each
__global__ void __launch_bounds__(256) sum_planes_kernel(const float4* __restrict__ part, int nplanes, size_t n4,
                                                         float4* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int k = 0; k < nplanes; ++k) {
            const float4 v = part[(size_t)k * n4 + i];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        out[i] = acc;
    }
}
// segments of the t sweep: enough blocks to fill the GPU a few times over (pure function of the geometry)
inline int red_tiles_per_seg(int nTb, int nUb, int nb) {
    long long blocks = (long long)nUb * nb;
    int nseg = (int)((RED_MIN_BLOCKS + blocks - 1) / blocks);
    if (nseg < 1) nseg = 1;
    if (nseg > nTb) nseg = nTb;
    return (nTb + nseg - 1) / nseg;
}
// zb[row, H .. H+7] = (1, 0, ..., 0): the ones column that turns the dW GEMM's extra output row into db
__global__ void __launch_bounds__(256) zb_ones_kernel(__nv_bfloat16* __restrict__ zb, size_t rows, int H, int zld) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) *reinterpret_cast<uint4*>(zb + r * zld + H) = make_uint4(0x00003F80u, 0u, 0u, 0u);   // bf16(1.0) = 0x3F80
}
// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}
// 2-D row-major bf16 matrix [rows x cols] (cols contiguous), box [box_rows x 64 cols], SWIZZLE_128B
inline bool make_tmap_bf16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                           uint32_t box_cols = 64) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// W^T (V,H) bf16 viewed as 3-D (k within a 64-wide K block, v, K block): one request brings `kblocks` K-major
// SWIZZLE_128B slabs [box_rows x 64] that land back to back in shared memory.
inline bool make_tmap_bf16_kblocks(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                                   uint32_t kblocks) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {64, rows, cols / 64};
    cuuint64_t strides[2] = {cols * 2, 128};
    cuuint32_t box[3] = {64, box_rows, kblocks};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 2-D row-major fp32 matrix [rows x cols], box [box_rows x box_cols]; swizzle128 requires box_cols*4 == 128
inline bool make_tmap_f32(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                          uint32_t box_cols, bool swizzle128) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 4};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct TcScratch {
    __nv_bfloat16 *Wt, *Wb, *dl, *zb, *dz;   // zb rows have stride tc_zld(H) (ones column at H); dz is bf16
    float* dWx;                             // (H+8, V) fp32: dW rows, then the db row produced by the ones column
    float *penc, *ppred;                    // partial planes of the single-pass reduction: (nUb, bchunk, maxT, H) and (nSeg, bchunk, maxU, H)
    float* gm;                              // (row blocks, V/32, 128) fp32: running maxima of the kept numerators (keep_activations)
    int* slot;                              // tile -> compact row block (per backward chunk)
    int* count;                             // number of valid tiles of the chunk
    int bchunk;          // utterances per backward pass
    size_t rows_chunk;   // bchunk * tiles_per_utt * 128
    size_t bytes;
};
inline TcScratch tc_scratch_layout(const rnntb200JointDesc& d, void* base) {
    TcScratch s{};
    const TcGeom g = tc_geometry(d.maxT, d.maxU, d.H, d.V);
    const size_t rows_utt = (size_t)g.nTb * g.nUb * 128;
    const size_t per_row = (size_t)d.V * 2 + (size_t)tc_zld(d.H) * 2 + (size_t)d.H * 2 + (size_t)(d.V / 32) * 4;
    const size_t budget = (size_t)16 << 30;
    size_t bc = budget / (rows_utt * per_row);
    if (bc < 1) bc = 1;
    if (bc > (size_t)d.B) bc = d.B;
    s.bchunk = (int)bc;
    s.rows_chunk = bc * rows_utt;
    char* p = static_cast<char*>(base);
    auto take = [&](size_t n) { char* r = p; p += (n + 255) / 256 * 256; return r; };
    s.Wt = reinterpret_cast<__nv_bfloat16*>(take((size_t)d.V * d.H * 2));
    s.Wb = reinterpret_cast<__nv_bfloat16*>(take((size_t)d.V * d.H * 2));
    s.dl = reinterpret_cast<__nv_bfloat16*>(take(s.rows_chunk * d.V * 2));
    s.zb = reinterpret_cast<__nv_bfloat16*>(take(s.rows_chunk * tc_zld(d.H) * 2));
    s.dz = reinterpret_cast<__nv_bfloat16*>(take(s.rows_chunk * d.H * 2));
    {
        const int tps = red_tiles_per_seg(g.nTb, g.nUb, (int)bc), nseg = (g.nTb + tps - 1) / tps;
        s.penc = reinterpret_cast<float*>(take((size_t)g.nUb * bc * d.maxT * d.H * 4));
        s.ppred = reinterpret_cast<float*>(take((size_t)nseg * bc * d.maxU * d.H * 4));
    }
    s.gm = reinterpret_cast<float*>(take(s.rows_chunk * (size_t)(d.V / 32) * 4));
    s.dWx = reinterpret_cast<float*>(take((size_t)(d.H + 8) * d.V * 4));
    s.slot = reinterpret_cast<int*>(take((size_t)bc * g.nTb * g.nUb * 4));
    s.count = reinterpret_cast<int*>(take(256));
    s.bytes = (size_t)(p - static_cast<char*>(base));
    return s;
}
inline size_t tc_scratch_bytes(const rnntb200JointDesc& d) {
    if (!tc_geometry(d.maxT, d.maxU, d.H, d.V).ok) return 0;
    return tc_scratch_layout(d, nullptr).bytes;
}

inline int tc_num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

inline bool tc_fill_params(const rnntb200JointDesc& d, const TcGeom& g, JointTcParams& p, const float* enc,
                           const float* pred, const float* bias, const int* labels, const int* ylen,
                           const int* xlen) {
    p = JointTcParams{};
    p.enc = enc; p.pred = pred; p.bias = bias; p.labels = labels; p.xlen = xlen; p.ylen = ylen;
    p.B = d.B; p.maxT = d.maxT; p.maxU = d.maxU; p.H = d.H; p.V = d.V; p.blank = d.blank_label;
    p.TT = g.TT; p.UU = g.UU; p.nTb = g.nTb; p.nUb = g.nUb; p.NC = g.NC; p.NCH = g.NCH; p.KB = g.KB;
    p.stages = g.stages;
    p.SK = (long long)(d.maxT + d.maxU - 1) * d.maxU;
    return true;
}

template <int MODE>
inline rnntStatus_t tc_launch(const TcGeom& g, const CUtensorMap& tm, const JointTcParams& p, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(joint_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)232448) != cudaSuccess)
            return RNNT_STATUS_EXECUTION_FAILED;
        attr_set = true;
    }
    const int ntiles = p.nb * p.nTb * p.nUb;
    const int grid = ntiles < tc_num_sms() ? ntiles : tc_num_sms();
    ScopedTimer tmr(MODE == 0 ? "joint_tc_kernel<fwd>" : "joint_tc_kernel<dlogits>", s);
    joint_tc_kernel<MODE><<<grid, TC_THREADS, g.smem_bytes, s>>>(tm, p);
    return cudaGetLastError() == cudaSuccess ? RNNT_STATUS_SUCCESS : RNNT_STATUS_EXECUTION_FAILED;
}

// Kernel generation: 1 = z resident in shared memory (joint_tc_kernel), 2 = z resident in tensor memory
// (joint_tc2_kernel).  RNNTB200_TC_VARIANT overrides the default for A/B measurements.
inline int tc_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RNNTB200_TC_VARIANT");
        v = e ? atoi(e) : RNNTB200_DEFAULT_TC_VARIANT;
        if (v < 1 || v > 3) v = RNNTB200_DEFAULT_TC_VARIANT;
    }
    return v;
}
inline int tc_dbg() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RNNTB200_DBG"); v = e ? atoi(e) : 0; }
    return v;
}
inline int tc_swap() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RNNTB200_TC2_SWAP"); v = e ? atoi(e) : 0; }
    return v;
}

inline rnntStatus_t tc_unsupported(const rnntb200JointDesc& d) {
    fprintf(stderr,
            "rnnt_b200: RNNTB200_BF16_TC needs H %% 64 == 0, V %% 64 == 0 and H <= 640 for the resident-z kernel "
            "(got H=%d V=%d); use RNNTB200_FP32_EXACT\n", d.H, d.V);
    return RNNT_STATUS_INVALID_VALUE;
}

template <int MODE>
inline rnntStatus_t tc_dispatch(const rnntb200JointDesc& d, const TcGeom& g, const TcScratch& sc, JointTcParams& p,
                                cudaStream_t s);

}  // namespace rb

#include "bwd_gemm.cuh"
#include "joint_tc2.cuh"
#include "joint_tc3.cuh"

namespace rb {
template <int MODE>
inline rnntStatus_t tc_dispatch(const rnntb200JointDesc& d, const TcGeom& g, const TcScratch& sc, JointTcParams& p,
                                cudaStream_t s) {
    CUtensorMap tm;
    const Tc2Geom g2 = tc2_geometry(d.H, d.V);
    const Tc2Geom g3 = tc3_geometry(d.H, d.V);
    if (tc_variant() == 3 && g3.ok) {
        CUtensorMap tmp, tme;
        if (!make_tmap_bf16_kblocks(&tm, sc.Wt, d.V, d.H, TC2_NC, g3.ks) ||
            !make_tmap_f32(&tmp, p.pred, (uint64_t)d.B * d.maxU, d.H, g.UU, 32, true) ||
            !make_tmap_f32(&tme, p.enc, (uint64_t)d.B * d.maxT, d.H, g.TT, 64, false))
            return RNNT_STATUS_EXECUTION_FAILED;
        p.NC = TC2_NC; p.NCH = d.V / TC2_NC; p.stages = g3.stages; p.nbuf = g3.nbuf; p.swap = 0; p.ks = g3.ks;
        p.dbg = tc_dbg();
        return tc3_launch<MODE>(g3, tm, tmp, tme, p, s);
    }
    if (tc_variant() == 2 && g2.ok) {
        if (!make_tmap_bf16_kblocks(&tm, sc.Wt, d.V, d.H, TC2_NC, g2.ks)) return RNNT_STATUS_EXECUTION_FAILED;
        p.NC = TC2_NC; p.NCH = d.V / TC2_NC; p.stages = g2.stages; p.nbuf = g2.nbuf; p.swap = tc_swap(); p.ks = g2.ks;
        p.dbg = tc_dbg();
        return tc2_launch<MODE>(g2, tm, p, s);
    }
    if (!make_tmap_bf16(&tm, sc.Wt, d.V, d.H, g.NC)) {
        fprintf(stderr, "rnnt_b200: cuTensorMapEncodeTiled failed\n");
        return RNNT_STATUS_EXECUTION_FAILED;
    }
    return tc_launch<MODE>(g, tm, p, s);
}
}  // namespace rb

namespace rb {

// pinned host word for the valid-tile count read-back (ragged batches)
inline int* host_count() {
    static int* h = nullptr;
    if (!h && cudaHostAlloc(reinterpret_cast<void**>(&h), sizeof(int), cudaHostAllocDefault) != cudaSuccess) h = nullptr;
    return h;
}
// slot[] / count of the tiles [b0, b0+nb) that intersect the valid lattice; returns the valid ROWS (or -1).  One 4-byte
// read-back + stream synchronise (before anything that depends on it is enqueued, so the GPU only idles for the
// launch latency).  Only used when the descriptor allows host synchronisation.
inline long long compact_tiles(const TcGeom& g, const TcScratch& sc, const int* xlen, const int* ylen, int b0, int ntiles,
                               cudaStream_t s) {
    tile_compact_kernel<<<1, 1024, 0, s>>>(xlen, ylen, b0, ntiles, g.nTb, g.nUb, g.TT, g.UU, sc.slot, sc.count);
    int* hcount = host_count();
    if (!hcount || cudaMemcpyAsync(hcount, sc.count, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
        return -1;
    return (long long)(*hcount) * 128;
}

// keep_activations: honoured when the generation-3 kernel runs and the whole batch is one workspace chunk
// (RNNTB200_KEEP=0 turns it off for A/B measurements).
inline bool tc_keep(const rnntb200JointDesc& d, const TcScratch& sc) {
    static int env = -1;
    if (env < 0) { const char* e = getenv("RNNTB200_KEEP"); env = e ? atoi(e) : 1; }
    return d.keep_activations && env && tc_variant() == 3 && tc3_geometry(d.H, d.V).ok && sc.bchunk >= d.B;
}
// Row count of the kept forward, remembered per workspace so that the backward does not have to read it back again
// (autograd may run the backward on another host thread: a mutex-protected table, not thread-local state).
struct KeptRows {
    std::mutex mu;
    std::unordered_map<const void*, long long> rows;
};
inline KeptRows& kept_rows() { static KeptRows k; return k; }

inline rnntStatus_t tc_forward(const rnntb200JointDesc& d, void* scratch, const float* enc, const float* pred,
                               const float* W, const float* bias, const int* labels, const int* ylen,
                               const int* xlen, float* lse, float* lpb, float* lpl, cudaStream_t s,
                               unsigned* launches) {
    const TcGeom g = tc_geometry(d.maxT, d.maxU, d.H, d.V);
    if (!g.ok) return tc_unsupported(d);
    TcScratch sc = tc_scratch_layout(d, scratch);
    convert_w_kernel<<<dim3((d.V + 31) / 32, (d.H + 31) / 32), 256, 0, s>>>(W, sc.Wt, sc.Wb, d.H, d.V);
    JointTcParams p;
    tc_fill_params(d, g, p, enc, pred, bias, labels, ylen, xlen);
    p.b0 = 0; p.nb = d.B;
    p.lse = lse; p.lpb = lpb; p.lpl = lpl;
    *launches += 2;
    if (!tc_keep(d, sc)) return tc_dispatch<0>(d, g, sc, p, s);
    // forward that keeps its activations: rows of dl / gm / zb are laid out exactly as the backward GEMMs want them
    const int ntiles = d.B * g.nTb * g.nUb;
    long long rows = (long long)ntiles * 128;
    if (d.allow_host_sync) {
        rows = compact_tiles(g, sc, xlen, ylen, 0, ntiles, s);
        if (rows < 0) return RNNT_STATUS_EXECUTION_FAILED;
        p.slot = sc.slot;
        *launches += 1;
    }
    {
        std::lock_guard<std::mutex> lk(kept_rows().mu);
        kept_rows().rows[scratch] = rows;
    }
    p.dl = sc.dl; p.gm = sc.gm; p.zb = sc.zb; p.zld = tc_zld(d.H);
    CUtensorMap tm, tmp, tme;
    const Tc2Geom g3 = tc3_geometry(d.H, d.V);
    if (!make_tmap_bf16_kblocks(&tm, sc.Wt, d.V, d.H, TC2_NC, g3.ks) ||
        !make_tmap_f32(&tmp, p.pred, (uint64_t)d.B * d.maxU, d.H, g.UU, 32, true) ||
        !make_tmap_f32(&tme, p.enc, (uint64_t)d.B * d.maxT, d.H, g.TT, 64, false))
        return RNNT_STATUS_EXECUTION_FAILED;
    p.NC = TC2_NC; p.NCH = d.V / TC2_NC; p.stages = g3.stages; p.nbuf = g3.nbuf; p.swap = 0; p.ks = g3.ks;
    p.dbg = tc_dbg();
    return tc3_launch<2>(g3, tm, tmp, tme, p, s);
}

// Library-owned side stream for the fork/join inside the backward (created once per process).
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    bool ok = false;
};
inline SideStream& side_stream() {
    static SideStream ss;
    if (!ss.ok) {
        ss.ok = cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking) == cudaSuccess &&
                cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) == cudaSuccess;
    }
    return ss;
}

// reduction phase: 0 = two streaming passes (g written back), 1 = single pass over whole tiles (RNNTB200_RED)
inline int red_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RNNTB200_RED"); v = e ? atoi(e) : 1; }
    return v;
}

inline rnntStatus_t tc_backward(const rnntb200JointDesc& d, void* scratch, const float* enc, const float* pred,
                                const float* W, const float* bias, const int* labels, const int* ylen,
                                const int* xlen, const float* lse, const float4* coef, float* d_enc, float* d_pred,
                                float* dW, float* db, cudaStream_t s, unsigned* launches, bool compact) {
    (void)W; (void)lse;
    const TcGeom g = tc_geometry(d.maxT, d.maxU, d.H, d.V);
    if (!g.ok) return tc_unsupported(d);
    TcScratch sc = tc_scratch_layout(d, scratch);   // Wt / Wb were produced by the forward call
    for (int b0 = 0; b0 < d.B; b0 += sc.bchunk) {
        const int nb = (d.B - b0 < sc.bchunk) ? d.B - b0 : sc.bchunk;
        JointTcParams p;
        tc_fill_params(d, g, p, enc, pred, bias, labels, ylen, xlen);
        p.b0 = b0; p.nb = nb; p.coef = coef; p.dl = sc.dl; p.zb = sc.zb; p.zld = tc_zld(d.H);
        const int ntiles = nb * g.nTb * g.nUb;
        size_t rows = (size_t)ntiles * 128;
        const int* slot = nullptr;
        const bool keep = tc_keep(d, sc);
        if (keep) {
            // the forward kept numerators / maxima / tanh outputs in exactly these rows: one streaming pass turns them into dlogits
            {
                std::lock_guard<std::mutex> lk(kept_rows().mu);
                auto it = kept_rows().rows.find(scratch);
                if (it == kept_rows().rows.end()) {
                    fprintf(stderr, "rnnt_b200: backward with keep_activations without a matching forward on this workspace\n");
                    return RNNT_STATUS_INVALID_VALUE;
                }
                rows = (size_t)it->second;
                kept_rows().rows.erase(it);
            }
            if (compact) slot = sc.slot;
            if (rows == 0) continue;
            p.slot = slot; p.gm = sc.gm;
            {
                const size_t gsm = (size_t)128 * (d.V / 32 + 1) * sizeof(float);   // the tile's maxima, transposed
                static size_t gsm_set = 0;
                if (gsm > 48 * 1024 && gsm > gsm_set) {
                    if (cudaFuncSetAttribute(dl_from_kept_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm) != cudaSuccess)
                        return RNNT_STATUS_EXECUTION_FAILED;
                    gsm_set = gsm;
                }
                ScopedTimer tmr("dl_from_kept_kernel", s);
                dl_from_kept_kernel<<<ntiles, 256, gsm, s>>>(p);
            }
            *launches += 1;
        } else {
            if (compact) {
                // Ragged batches: only tiles that intersect the valid lattice get rows in dl / zb / dZ, so the two GEMMs
                // run over the valid rows instead of the padded ones (off by default in the C ABI: allow_host_sync == 0).
                const long long r = compact_tiles(g, sc, xlen, ylen, b0, ntiles, s);
                if (r < 0) return RNNT_STATUS_EXECUTION_FAILED;
                rows = (size_t)r;
                slot = sc.slot;
                *launches += 1;
                if (rows == 0) continue;
            }
            p.slot = slot;
            rnntStatus_t st1 = tc_dispatch<1>(d, g, sc, p, s);
            if (st1) return st1;
        }
        const RowMap m{g.TT, g.UU, g.nTb, g.nUb, b0, ilog2(g.UU), slot};
        rnntStatus_t st;
        // (the generation-3 kernel writes the ones column of zb itself)
        if (!(tc_variant() == 3 && tc3_geometry(d.H, d.V).ok))
            zb_ones_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(sc.zb, rows, d.H, tc_zld(d.H));
        // dZ[rows,H] (bf16) = dl[rows,V] . Wb[H,V]^T, then two independent branches:
        //   side stream : g = dZ*sech^2 -> d_enc, d_pred        (memory / MUFU bound)
        //   main stream : dWx[H+8,V] (+)= zb^T . dl  (row H = db) (tensor bound)
        st = bwd_gemm_dz(d, sc, rows, s, launches);
        if (st) return st;
        SideStream& ss = side_stream();
        if (!ss.ok) return RNNT_STATUS_EXECUTION_FAILED;
        cudaEventRecord(ss.fork, s);
        cudaStreamWaitEvent(ss.stream, ss.fork, 0);
        {
            if (red_variant() == 1 && g.UU == RED_UU && d.H <= 8 * RED_THREADS && d.H % 8 == 0) {
                // nb (not bchunk) utterances in this launch: the planes are laid out for nb
                const int tps = red_tiles_per_seg(g.nTb, g.nUb, nb), nseg = (g.nTb + tps - 1) / tps;
                const dim3 grid(g.nUb, nseg, nb);
                ScopedTimer* t1 = new ScopedTimer("dencpred_tiles_kernel", ss.stream);
                dencpred_tiles_kernel<<<grid, RED_THREADS, 0, ss.stream>>>(sc.dz, enc, pred, xlen, ylen, m, d.maxT, d.maxU, d.H, tps,
                                                                       sc.penc, sc.ppred);
                delete t1; t1 = new ScopedTimer("sum_planes_kernel", ss.stream);
                const size_t ne4 = (size_t)nb * d.maxT * d.H / 4, np4 = (size_t)nb * d.maxU * d.H / 4;
                sum_planes_kernel<<<(unsigned)((ne4 + 255) / 256 < 4096 ? (ne4 + 255) / 256 : 4096), 256, 0, ss.stream>>>(
                    reinterpret_cast<const float4*>(sc.penc), g.nUb, ne4, reinterpret_cast<float4*>(d_enc + (size_t)b0 * d.maxT * d.H));
                sum_planes_kernel<<<(unsigned)((np4 + 255) / 256 < 4096 ? (np4 + 255) / 256 : 4096), 256, 0, ss.stream>>>(
                    reinterpret_cast<const float4*>(sc.ppred), nseg, np4, reinterpret_cast<float4*>(d_pred + (size_t)b0 * d.maxU * d.H));
                delete t1;
            } else {
                const int rthreads = ((d.H / 8 + 31) / 32) * 32;
                ScopedTimer* t1 = new ScopedTimer("denc_rows_kernel", ss.stream);
                denc_rows_kernel<<<dim3(d.maxT, nb), rthreads, 0, ss.stream>>>(sc.dz, enc, pred, xlen, ylen, m, d.maxT, d.maxU, d.H, d_enc);
                delete t1; t1 = new ScopedTimer("dpred_rows_kernel", ss.stream);
                dpred_rows_kernel<<<dim3(d.maxU, nb), rthreads, 0, ss.stream>>>(sc.dz, xlen, ylen, m, d.maxU, d.H, d_pred);
                delete t1;
            }
        }
        cudaEventRecord(ss.join, ss.stream);
        st = bwd_gemm_dw(d, sc, rows, /*accumulate=*/b0 > 0, s, launches);
        if (st) return st;
        cudaStreamWaitEvent(s, ss.join, 0);
        *launches += 3;
        if (cudaGetLastError() != cudaSuccess) return RNNT_STATUS_EXECUTION_FAILED;
    }
    if (cudaMemcpyAsync(dW, sc.dWx, sizeof(float) * (size_t)d.H * d.V, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(db, sc.dWx + (size_t)d.H * d.V, sizeof(float) * d.V, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
        return RNNT_STATUS_MEMOPS_FAILED;
    return RNNT_STATUS_SUCCESS;
}

}  // namespace rb
