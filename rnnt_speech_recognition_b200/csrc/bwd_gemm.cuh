// bwd_gemm.cuh -- the two plain GEMMs of the bf16 backward over the materialised bf16 dlogits rows:
//     dZ[rows,H]    = dl[rows,V] . Wb[H,V]^T            (bf16 out: halves the traffic of the two reduction passes)
//     dWx[H+8,V] (+)= zb[rows,0:H+8]^T . dl[rows,V]     (fp32 out, accumulated across utterance chunks; zb carries a
//                                                        ones column at index H, so row H of dWx is db = sum_rows dl)
// RNNTB200_BWD_CUBLAS: interim library path (cuBLAS is allowed for PLAIN GEMMs; these two have no
// fused prologue/epilogue).  It is the baseline the hand-written tcgen05 kernels are checked against.
#pragma once
#include <cublas_v2.h>

namespace rb {

inline cublasHandle_t tc_cublas() {
    static cublasHandle_t h = nullptr;
    if (!h && cublasCreate(&h) != CUBLAS_STATUS_SUCCESS) h = nullptr;
    return h;
}

inline rnntStatus_t bwd_gemm_dz(const rnntb200JointDesc& d, const TcScratch& sc, size_t rows, cudaStream_t s,
                                unsigned* launches) {
    cublasHandle_t h = tc_cublas();
    if (!h || cublasSetStream(h, s) != CUBLAS_STATUS_SUCCESS) return RNNT_STATUS_EXECUTION_FAILED;
    const float one = 1.f, zero = 0.f;
    ScopedTimer t("gemm dZ=dl.W^T (cublas)", s);
    // row-major dZ[rows,H] == column-major [H,rows] = Wb_cm[V,H]^T . dl_cm[V,rows]
    if (cublasGemmEx(h, CUBLAS_OP_T, CUBLAS_OP_N, d.H, (int)rows, d.V, &one, sc.Wb, CUDA_R_16BF, d.V, sc.dl,
                     CUDA_R_16BF, d.V, &zero, sc.dz, CUDA_R_16BF, d.H, CUBLAS_COMPUTE_32F,
                     CUBLAS_GEMM_DEFAULT) != CUBLAS_STATUS_SUCCESS)
        return RNNT_STATUS_EXECUTION_FAILED;
    (void)launches;   // library GEMMs are not counted in rnntb200_launch_count (own kernels only)
    return RNNT_STATUS_SUCCESS;
}

inline rnntStatus_t bwd_gemm_dw(const rnntb200JointDesc& d, const TcScratch& sc, size_t rows, bool accumulate,
                                cudaStream_t s, unsigned* launches) {
    cublasHandle_t h = tc_cublas();
    if (!h || cublasSetStream(h, s) != CUBLAS_STATUS_SUCCESS) return RNNT_STATUS_EXECUTION_FAILED;
    const float one = 1.f, beta = accumulate ? 1.f : 0.f;
    ScopedTimer t("gemm dW=z^T.dl (cublas)", s);
    // row-major dWx[H+8,V] == column-major [V,H+8] = dl_cm[V,rows] . zb_cm[H+8,rows]^T   (zb_cm has leading dimension tc_zld(H))
    if (cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_T, d.V, d.H + 8, (int)rows, &one, sc.dl, CUDA_R_16BF, d.V, sc.zb,
                     CUDA_R_16BF, tc_zld(d.H), &beta, sc.dWx, CUDA_R_32F, d.V, CUBLAS_COMPUTE_32F,
                     CUBLAS_GEMM_DEFAULT) != CUBLAS_STATUS_SUCCESS)
        return RNNT_STATUS_EXECUTION_FAILED;
    (void)launches;   // library GEMMs are not counted in rnntb200_launch_count (own kernels only)
    return RNNT_STATUS_SUCCESS;
}

}  // namespace rb
