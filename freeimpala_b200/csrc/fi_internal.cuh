// Internal (non-ABI) entry points shared between the translation units of the library.
#pragma once
#include "fi_common.cuh"

namespace fi {

int launch_gather(const void* ring_base, size_t capacity, size_t slot_bytes, size_t first, size_t m, void* dst,
                  cudaStream_t stream);
int launch_opt(int opt_kind, double lr, int64_t step, size_t n, float* p, const float* g, float* m, float* v,
               float grad_scale, cudaStream_t stream);
int launch_vtrace_scan(int m, int t, const float* log_rho, const float* discount, const float* reward,
                       const float* value, const float* bootstrap, float rho_bar, float c_bar, float pg_rho_bar,
                       float lambda_, float* vs, float* pg_adv, cudaStream_t stream);
int launch_vtrace_loss_head(const void* batch, int m, int t, const float* head, int ldh, float rho_bar, float c_bar,
                            float pg_rho_bar, float lambda_, float baseline_cost, float entropy_cost, float* dhead,
                            float* vs, float* pg_adv, double* losses, cudaStream_t stream,
                            float* dhead_hi = nullptr, float* dhead_lo = nullptr, int ld_split = 0);

// trans: 0 "NT" C = A[m,k] B[n,k]^T; 1 "NN" C = A[m,k] B[k,n]; 2 "TN" C = A[k,m]^T B[k,n].
int launch_gemm_simt(int trans, int m, int n, int k, const float* a, int lda, const float* b, int ldb, float* c,
                     int ldc, const float* bias, int relu, const float* mask, int ldmask, void* workspace,
                     size_t workspace_bytes, cudaStream_t stream);
size_t gemm_simt_workspace_bytes(int trans, int m, int n, int k);
int launch_colsum(const float* x, int ldx, int m, int n, float* out, void* workspace, size_t workspace_bytes,
                  cudaStream_t stream);
int launch_colsum2(const float* x, const float* x2, int ldx, int m, int n, float* out, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream);
size_t colsum_workspace_bytes(int m, int n);

// out[i] = sum_s partial[s * stride + i] in fixed order (deterministic split-K reduction).
int launch_reduce_splits(const float* partial, int splits, size_t stride, size_t n, float* out, cudaStream_t stream);

// ---- tcgen05 3xTF32 path (gemm_tc.cu) -----------------------------------------------------------
// An fp32 matrix carried as the exact pair x = hi + lo (hi: low 13 mantissa bits cleared), row stride ld.
struct SplitMat {
    const float* hi;
    const float* lo;
    int ld;
};
struct TcOut {
    float* c; int ldc;                       // plain fp32 output (may be null)
    float* c_hi; float* c_lo; int ld_split;  // hi/lo output for the next GEMM (may be null)
    int transpose;                           // plain output written as c[col * ldc + row]
    const uint32_t* mask_bits_in;            // bit-packed ReLU decisions applied to the output (backward), or null
    uint32_t* mask_bits_out;                 // bit-packed (output > 0) written by a ReLU epilogue (forward), or null
    int mask_ldw;                            // words per mask row (a multiple of 4)
    float* colsum_out;                       // [4 * ceil(m/128), n] per-32-row column sums of the output (split-output path), or null
};
int launch_split_tf32(const float* x, int ld_in, size_t rows, int cols, int ld_out, float* hi, float* lo, cudaStream_t st);
int launch_gemm_tc_split(int trans, int m, int n, int k, SplitMat a, SplitMat b, TcOut out, const float* bias, int relu,
                         const float* mask, int ldmask, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t gemm_tc_split_workspace_bytes(int trans, int m, int n, int k);
bool gemm_tc_available();

// Dispatching GEMM (gemm.cu): picks the tcgen05 3xTF32 kernel or the SIMT kernel per fi_gemm_mode.
int launch_gemm(int mode, int trans, int m, int n, int k, const float* a, int lda, const float* b, int ldb, float* c,
                int ldc, const float* bias, int relu, const float* mask, int ldmask, void* workspace,
                size_t workspace_bytes, cudaStream_t stream);
size_t gemm_workspace_bytes(int mode, int trans, int m, int n, int k);

}  // namespace fi
