// GEMM dispatch: tcgen05 3xTF32 (gemm_tc.cu) where the shape allows, else fp32 SIMT.
#include "fi_internal.cuh"

namespace fi {

// gemm_tc.cu
bool gemm_tc_supported(int trans, int m, int n, int k, const float* a, int lda, const float* b, int ldb,
                       const float* c, int ldc);
int launch_gemm_tc(int trans, int m, int n, int k, const float* a, int lda, const float* b, int ldb, float* c,
                   int ldc, const float* bias, int relu, const float* mask, int ldmask, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream);
size_t gemm_tc_workspace_bytes(int trans, int m, int n, int k);
int launch_gemm_h(int trans, int m, int n, int k, const float* a, int lda, const float* b, int ldb, float* c,
                  int ldc, const float* bias, int relu, const float* mask, int ldmask, void* workspace,
                  size_t workspace_bytes, cudaStream_t stream);
size_t gemm_h_workspace_bytes(int trans, int m, int n, int k);

size_t gemm_workspace_bytes(int mode, int trans, int m, int n, int k) {
    size_t a = gemm_simt_workspace_bytes(trans, m, n, k);
    size_t b = mode == FI_GEMM_SIMT ? 0 : gemm_tc_workspace_bytes(trans, m, n, k);
    if (mode == FI_GEMM_TCGEN05_F16) b = gemm_h_workspace_bytes(trans, m, n, k);
    return a > b ? a : b;
}

int launch_gemm(int mode, int trans, int m, int n, int k, const float* a, int lda, const float* b, int ldb, float* c,
                int ldc, const float* bias, int relu, const float* mask, int ldmask, void* workspace,
                size_t workspace_bytes, cudaStream_t stream) {
    if (mode == FI_GEMM_TCGEN05_F16) {
        if (!gemm_tc_supported(trans, m, n, k, a, lda, b, ldb, c, ldc))
            return set_error(FI_ERR_ARG, "gemm: shape m=%d n=%d k=%d trans=%d cannot use the tcgen05 path", m, n, k, trans);
        return launch_gemm_h(trans, m, n, k, a, lda, b, ldb, c, ldc, bias, relu, mask, ldmask, workspace, workspace_bytes, stream);
    }
    if (mode != FI_GEMM_SIMT) {
        // AUTO: the tensor-core path pays a split pre-pass for plain fp32 operands; use it from ~64 MFLOP up. (Measured in
        // round 2: sending the FarmerLstm dense stack at batch 1024 -- 0.5 GFLOP per product -- to the fp32 FFMA kernel instead
        // costs 56 us per product against ~25 us for split + tensor-core product: the FFMA kernel is inefficient at m = 1024.)
        const bool big = 2.0 * (double)m * (double)n * (double)k >= 64e6;
        if ((mode == FI_GEMM_TCGEN05 || big) && gemm_tc_supported(trans, m, n, k, a, lda, b, ldb, c, ldc) &&
            (mode == FI_GEMM_TCGEN05 || (workspace && workspace_bytes >= gemm_tc_workspace_bytes(trans, m, n, k))))
            return launch_gemm_tc(trans, m, n, k, a, lda, b, ldb, c, ldc, bias, relu, mask, ldmask, workspace,
                                  workspace_bytes, stream);
        if (mode == FI_GEMM_TCGEN05)
            return set_error(FI_ERR_ARG, "gemm: shape m=%d n=%d k=%d trans=%d cannot use the tcgen05 path", m, n, k, trans);
    }
    return launch_gemm_simt(trans, m, n, k, a, lda, b, ldb, c, ldc, bias, relu, mask, ldmask, workspace,
                            workspace_bytes, stream);
}

}  // namespace fi

extern "C" int fi_op_gemm(int trans, int m, int n, int k, const float* a, int lda, const float* b, int ldb, float* c,
                          int ldc, const float* bias, int relu, int mode, void* workspace, size_t workspace_bytes,
                          void* stream) {
    return fi::launch_gemm(mode, trans, m, n, k, a, lda, b, ldb, c, ldc, bias, relu, nullptr, 0, workspace,
                           workspace_bytes, (cudaStream_t)stream);
}
extern "C" size_t fi_op_gemm_workspace_bytes(int trans, int m, int n, int k, int mode) {
    return fi::gemm_workspace_bytes(mode, trans, m, n, k);
}
