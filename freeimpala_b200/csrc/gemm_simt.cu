// fp32 SIMT GEMM family (FFMA), used where exact fp32 products are required or a shape does not
// fit the tcgen05 path (e.g. the 17-row policy/value head, K = 162). One templated kernel covers
// the three products of a Linear layer (reference cmd/libtorch_bench/main.cpp:17-22,31-36 runs
// them through libtorch's addmm):
//   forward  Y[m,n]  = X[m,k] W[n,k]^T + b (ReLU)        A k-contiguous, B k-contiguous
//   dgrad    dX[m,k] = dY[m,n] W[n,k]  (* relu'(X))      A k-contiguous, B n-contiguous
//   wgrad    dW[n,k] = dY[m,n]^T X[m,k]                  A m-contiguous, B m-contiguous, split-K
// CTA tile 128 x BN x 16, 256 threads, 8 x (BN/16) register tile, double-buffered shared memory.
#include "fi_internal.cuh"

namespace fi {

constexpr int kBM = 128, kBK = 16, kGemmThreads = 256;

struct GemmArgs {
    const float* a; int lda;
    const float* b; int ldb;
    float* c; int ldc;
    int m, n, k;
    const float* bias;      // [n] or null
    const float* mask;      // [m, ldmask]: c = mask > 0 ? c : 0 (ReLU backward) or null
    int ldmask;
    int relu;
    int k_per_split;        // k range per blockIdx.z
    size_t split_stride;    // elements between split slabs of c (split-K partials)
};

// A(i,kk): ATRANS ? a[kk*lda + i] : a[i*lda + kk].   B(kk,j): BTRANS ? b[kk*ldb + j] : b[j*ldb + kk].
template <int BN, bool ATRANS, bool BTRANS>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_simt_kernel(GemmArgs g) {
    constexpr int TN = BN / 16;       // columns per thread
    constexpr int PAD = 4;
    __shared__ __align__(16) float As[2][kBK][kBM + PAD];
    __shared__ __align__(16) float Bs[2][kBK][BN + PAD];
    const int tid = threadIdx.x;
    const int bm0 = blockIdx.y * kBM, bn0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * g.k_per_split;
    const int k_end = min(g.k, k_begin + g.k_per_split);
    const float* __restrict__ A = g.a;
    const float* __restrict__ B = g.b;

    // ---- global -> register staging --------------------------------------------------
    // A tile: 128 x 16 floats = 512 float4; B tile: BN x 16 floats = BN*4 float4.
    constexpr int A_V = (kBM * kBK / 4) / kGemmThreads;                 // 2
    constexpr int B_V = (BN * kBK / 4 + kGemmThreads - 1) / kGemmThreads;  // 2 (BN=128) or 1 (BN=32, half idle)
    float4 ra[A_V], rb[B_V];

    auto load_a = [&](int k0) {
#pragma unroll
        for (int v = 0; v < A_V; v++) {
            const int idx = tid + v * kGemmThreads;
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (!ATRANS) {        // k-contiguous: idx -> (row, kq)
                const int row = idx >> 2, kq = (idx & 3) << 2;
                const int gi = bm0 + row, gk = k0 + kq;
                if (gi < g.m) {
                    const float* p = A + (size_t)gi * g.lda + gk;
                    if (gk + 3 < k_end && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) r = __ldg(reinterpret_cast<const float4*>(p));
                    else {
                        if (gk + 0 < k_end) r.x = __ldg(p + 0);
                        if (gk + 1 < k_end) r.y = __ldg(p + 1);
                        if (gk + 2 < k_end) r.z = __ldg(p + 2);
                        if (gk + 3 < k_end) r.w = __ldg(p + 3);
                    }
                }
            } else {                         // m-contiguous: idx -> (kk, mq)
                const int kk = idx >> 5, mq = (idx & 31) << 2;
                const int gk = k0 + kk, gi = bm0 + mq;
                if (gk < k_end) {
                    const float* p = A + (size_t)gk * g.lda + gi;
                    if (gi + 3 < g.m && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) r = __ldg(reinterpret_cast<const float4*>(p));
                    else {
                        if (gi + 0 < g.m) r.x = __ldg(p + 0);
                        if (gi + 1 < g.m) r.y = __ldg(p + 1);
                        if (gi + 2 < g.m) r.z = __ldg(p + 2);
                        if (gi + 3 < g.m) r.w = __ldg(p + 3);
                    }
                }
            }
            ra[v] = r;
        }
    };
    auto load_b = [&](int k0) {
#pragma unroll
        for (int v = 0; v < B_V; v++) {
            const int idx = tid + v * kGemmThreads;
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < BN * kBK / 4) {
                if constexpr (!BTRANS) {    // k-contiguous: idx -> (col, kq)
                    const int col = idx >> 2, kq = (idx & 3) << 2;
                    const int gj = bn0 + col, gk = k0 + kq;
                    if (gj < g.n) {
                        const float* p = B + (size_t)gj * g.ldb + gk;
                        if (gk + 3 < k_end && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) r = __ldg(reinterpret_cast<const float4*>(p));
                        else {
                            if (gk + 0 < k_end) r.x = __ldg(p + 0);
                            if (gk + 1 < k_end) r.y = __ldg(p + 1);
                            if (gk + 2 < k_end) r.z = __ldg(p + 2);
                            if (gk + 3 < k_end) r.w = __ldg(p + 3);
                        }
                    }
                } else {                     // n-contiguous: idx -> (kk, nq)
                    constexpr int NQ = BN / 4;
                    const int kk = idx / NQ, nq = (idx % NQ) << 2;
                    const int gk = k0 + kk, gj = bn0 + nq;
                    if (gk < k_end) {
                        const float* p = B + (size_t)gk * g.ldb + gj;
                        if (gj + 3 < g.n && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) r = __ldg(reinterpret_cast<const float4*>(p));
                        else {
                            if (gj + 0 < g.n) r.x = __ldg(p + 0);
                            if (gj + 1 < g.n) r.y = __ldg(p + 1);
                            if (gj + 2 < g.n) r.z = __ldg(p + 2);
                            if (gj + 3 < g.n) r.w = __ldg(p + 3);
                        }
                    }
                }
            }
            rb[v] = r;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int v = 0; v < A_V; v++) {
            const int idx = tid + v * kGemmThreads;
            if constexpr (!ATRANS) {
                const int row = idx >> 2, kq = (idx & 3) << 2;
                As[buf][kq + 0][row] = ra[v].x; As[buf][kq + 1][row] = ra[v].y;
                As[buf][kq + 2][row] = ra[v].z; As[buf][kq + 3][row] = ra[v].w;
            } else {
                const int kk = idx >> 5, mq = (idx & 31) << 2;
                *reinterpret_cast<float4*>(&As[buf][kk][mq]) = ra[v];
            }
        }
#pragma unroll
        for (int v = 0; v < B_V; v++) {
            const int idx = tid + v * kGemmThreads;
            if (idx < BN * kBK / 4) {
                if constexpr (!BTRANS) {
                    const int col = idx >> 2, kq = (idx & 3) << 2;
                    Bs[buf][kq + 0][col] = rb[v].x; Bs[buf][kq + 1][col] = rb[v].y;
                    Bs[buf][kq + 2][col] = rb[v].z; Bs[buf][kq + 3][col] = rb[v].w;
                } else {
                    constexpr int NQ = BN / 4;
                    const int kk = idx / NQ, nq = (idx % NQ) << 2;
                    *reinterpret_cast<float4*>(&Bs[buf][kk][nq]) = rb[v];
                }
            }
        }
    };

    // ---- main loop ----------------------------------------------------------------------
    const int ty = tid >> 4, tx = tid & 15;   // 16 x 16 threads
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

    const int nk = (k_end - k_begin + kBK - 1) / kBK;
    if (nk > 0) {
        load_a(k_begin);
        load_b(k_begin);
        store_tiles(0);
        __syncthreads();
    }
    for (int it = 0; it < nk; it++) {
        const int buf = it & 1;
        if (it + 1 < nk) {
            load_a(k_begin + (it + 1) * kBK);
            load_b(k_begin + (it + 1) * kBK);
        }
#pragma unroll
        for (int kk = 0; kk < kBK; kk++) {
            float af[8], bf[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            af[0] = a0.x; af[1] = a0.y; af[2] = a0.z; af[3] = a0.w;
            af[4] = a1.x; af[5] = a1.y; af[6] = a1.z; af[7] = a1.w;
            if constexpr (TN == 8) {
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
                bf[0] = b0.x; bf[1] = b0.y; bf[2] = b0.z; bf[3] = b0.w;
                bf[4] = b1.x; bf[5] = b1.y; bf[6] = b1.z; bf[7] = b1.w;
            } else {
#pragma unroll
                for (int j = 0; j < TN; j++) bf[j] = Bs[buf][kk][tx * TN + j];
            }
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
        }
        if (it + 1 < nk) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue -------------------------------------------------------------------------
    float* C = g.c + (size_t)blockIdx.z * g.split_stride;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int gi = bm0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gi >= g.m) continue;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            int gj;
            if constexpr (TN == 8) gj = bn0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            else gj = bn0 + tx * TN + j;
            if (gj >= g.n) continue;
            float v = acc[i][j];
            if (g.bias) v += __ldg(g.bias + gj);
            if (g.relu) v = fmaxf(v, 0.f);
            if (g.mask) v = (__ldg(g.mask + (size_t)gi * g.ldmask + gj) > 0.f) ? v : 0.f;
            C[(size_t)gi * g.ldc + gj] = v;
        }
    }
}

// out[i] = sum_s partial[s * stride + i], fixed order (deterministic split-K reduction).
__global__ void reduce_splits_kernel(const float* __restrict__ partial, int splits, size_t stride, size_t n,
                                     float* __restrict__ out) {
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < splits; k++) s += __ldg(partial + (size_t)k * stride + i);
        out[i] = s;
    }
}

// Column sums of dY[m, n] (bias gradient): partial[z][j] = sum over a row range, then reduced.
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ x, const float* __restrict__ x2, int ldx, int m, int n, int rows_per_block,
                      float* __restrict__ partial) {
    // block handles rows [r0, r1) and 32 columns (blockIdx.x); 8 row-lanes x 32 column-lanes
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + cx;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(m, r0 + rows_per_block);
    float s = 0.f;
    if (col < n)
        for (int r = r0 + ry; r < r1; r += 8) {
            s += __ldg(x + (size_t)r * ldx + col);
            if (x2) s += __ldg(x2 + (size_t)r * ldx + col);  // hi/lo pair: the column sum of hi + lo
        }
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && col < n) {
        float tsum = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) tsum += red[i][cx];
        partial[(size_t)blockIdx.y * n + col] = tsum;
    }
}

int launch_reduce_splits(const float* partial, int splits, size_t stride, size_t n, float* out, cudaStream_t stream) {
    int blocks = (int)((n + 255) / 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    LaunchScope lr("reduce_splits_kernel", stream, 4.0 * (double)n * (splits + 1), kWorkBytes);
    launch_pdl(reduce_splits_kernel, dim3(blocks), dim3(256), 0, stream, partial, splits, stride, n, out);
    return lr.done();
}

static int pick_splits(int tiles, int k, int max_splits) {
    // enough CTAs for ~2 waves of 148 SMs, at least 256 reduction steps per split
    int want = (2 * kNumSMs + tiles - 1) / tiles;
    int by_k = k / 256;
    if (by_k < 1) by_k = 1;
    int s = want < by_k ? want : by_k;
    if (s > max_splits) s = max_splits;
    return s < 1 ? 1 : s;
}

size_t gemm_simt_workspace_bytes(int trans, int m, int n, int k) {
    if (trans != 2) return 0;
    const int tiles = ((m + kBM - 1) / kBM) * ((n + 127) / 128);
    const int splits = pick_splits(tiles, k, 64);
    return splits > 1 ? (size_t)splits * m * n * sizeof(float) : 0;
}

// trans: 0 "NT" A[m,k] B[n,k]; 1 "NN" A[m,k] B[k,n]; 2 "TN" A[k,m] B[k,n] (split-K, workspace).
int launch_gemm_simt(int trans, int m, int n, int k, const float* a, int lda, const float* b, int ldb, float* c,
                     int ldc, const float* bias, int relu, const float* mask, int ldmask, void* workspace,
                     size_t workspace_bytes, cudaStream_t stream) {
    if (m <= 0 || n <= 0) return FI_OK;
    if (!a || !b || !c || k < 0) return set_error(FI_ERR_ARG, "gemm: bad argument");
    GemmArgs g;
    g.a = a; g.lda = lda; g.b = b; g.ldb = ldb; g.c = c; g.ldc = ldc;
    g.m = m; g.n = n; g.k = k; g.bias = bias; g.mask = mask; g.ldmask = ldmask; g.relu = relu;
    g.k_per_split = k; g.split_stride = 0;
    const bool narrow = n <= 32;
    const int bn = narrow ? 32 : 128;
    dim3 grid((n + bn - 1) / bn, (m + kBM - 1) / kBM, 1);
    int splits = 1;
    if (trans == 2) {
        splits = pick_splits(grid.x * grid.y, k, 64);
        if (splits > 1) {
            const size_t need = (size_t)splits * m * n * sizeof(float);
            if (!workspace || workspace_bytes < need) {
                splits = 1;  // no workspace: fall back to one (slow) split rather than fail
            } else {
                if (bias || relu || mask) return set_error(FI_ERR_ARG, "gemm TN: no epilogue with split-K");
                g.k_per_split = ((k + splits - 1) / splits + kBK - 1) / kBK * kBK;
                splits = (k + g.k_per_split - 1) / g.k_per_split;
                g.c = (float*)workspace; g.ldc = n; g.split_stride = (size_t)m * n;
                grid.z = splits;
            }
        }
    }
    if (trans < 0 || trans > 2) return set_error(FI_ERR_ARG, "gemm: trans must be 0, 1 or 2");
    LaunchScope ls("gemm_simt_kernel", stream, 2.0 * (double)m * (double)n * (double)k, kWorkFlops);
#define FI_LAUNCH(BN, AT, BT) gemm_simt_kernel<BN, AT, BT><<<grid, kGemmThreads, 0, stream>>>(g)
    if (trans == 0) { if (narrow) FI_LAUNCH(32, false, false); else FI_LAUNCH(128, false, false); }
    else if (trans == 1) { if (narrow) FI_LAUNCH(32, false, true); else FI_LAUNCH(128, false, true); }
    else { if (narrow) FI_LAUNCH(32, true, true); else FI_LAUNCH(128, true, true); }
#undef FI_LAUNCH
    FI_TRY(ls.done());
    if (trans == 2 && splits > 1) {
        const size_t total = (size_t)m * n;
        if (ldc != n) return set_error(FI_ERR_ARG, "gemm TN split-K: ldc must equal n");
        int blocks = (int)((total + 255) / 256);
        if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        LaunchScope lr("reduce_splits_kernel", stream, 4.0 * (double)total * (splits + 1), kWorkBytes);
        launch_pdl(reduce_splits_kernel, dim3(blocks), dim3(256), 0, stream, (const float*)workspace, splits, total, total, c);
        FI_TRY(lr.done());
    }
    return FI_OK;
}

// out[n] = column sums of x[m, n] (ldx). workspace >= colsum_workspace_bytes(m, n).
// Two deterministic levels: row blocks of ~128 rows (8 row lanes x 16 rows, 32 columns per CTA), then one warp per
// column adds the block partials with a shuffle tree. (The first version used row blocks of 512+ rows and a serial
// loop over the partials in 1-2 CTAs: 20 us of dependent loads for 6.5 MB.)
static int colsum_row_blocks(int m) {
    int rb = (m + 127) / 128;
    if (rb > 1024) rb = 1024;
    return rb < 1 ? 1 : rb;
}
size_t colsum_workspace_bytes(int m, int n) { return (size_t)colsum_row_blocks(m) * n * sizeof(float); }

__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ partial, int rb, int n, float* __restrict__ out) {
    const int col = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (col >= n) return;
    float s = 0.f;
    for (int p = lane; p < rb; p += 32) s += __ldg(partial + (size_t)p * n + col);
    s = warp_sum(s);
    if (lane == 0) out[col] = s;
}

int launch_colsum(const float* x, int ldx, int m, int n, float* out, void* workspace, size_t workspace_bytes,
                  cudaStream_t stream) {
    return launch_colsum2(x, nullptr, ldx, m, n, out, workspace, workspace_bytes, stream);
}
int launch_colsum2(const float* x, const float* x2, int ldx, int m, int n, float* out, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream) {
    if (n <= 0) return FI_OK;
    const int rb = colsum_row_blocks(m);
    const int rows_per_block = (m + rb - 1) / rb;
    if (!workspace || workspace_bytes < (size_t)rb * n * sizeof(float)) return set_error(FI_ERR_ARG, "colsum: workspace too small");
    dim3 grid((n + 31) / 32, rb);
    LaunchScope lc("colsum_partial_kernel", stream, (x2 ? 8.0 : 4.0) * (double)m * n, kWorkBytes);
    colsum_partial_kernel<<<grid, 256, 0, stream>>>(x, x2, ldx, m, n, rows_per_block, (float*)workspace);
    FI_TRY(lc.done());
    LaunchScope lr("colsum_final_kernel", stream, 4.0 * (double)n * (rb + 1), kWorkBytes);
    colsum_final_kernel<<<(n + 7) / 8, 256, 0, stream>>>((const float*)workspace, rb, n, out);
    return lr.done();
}

// Zero up to two small device buffers (word counts) with one tiny kernel: a cudaMemsetAsync may go through a copy engine,
// and the learner step keeps its stream free of copy-engine hand-overs (see OptExtras in adam.cu).
__global__ void zero_words_kernel(uint32_t* a, int na, uint32_t* b, int nb) {
    pdl_wait();
    for (int i = threadIdx.x; i < na; i += blockDim.x) a[i] = 0u;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) b[i] = 0u;
}
__global__ void copy_words_kernel(float* __restrict__ dst, const float* __restrict__ src, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
int launch_copy_words(float* dst, const float* src, int n, cudaStream_t st) {
    if (n <= 0) return FI_OK;
    LaunchScope ls("copy_words_kernel", st, 8.0 * n, kWorkBytes);
    copy_words_kernel<<<(n + 255) / 256 > 64 ? 64 : (n + 255) / 256, 256, 0, st>>>(dst, src, n);
    return ls.done();
}
int launch_zero2(void* a, size_t a_bytes, void* b, size_t b_bytes, cudaStream_t st) {
    if ((a_bytes | b_bytes) & 3) return set_error(FI_ERR_ARG, "launch_zero2: sizes must be multiples of 4");
    LaunchScope ls("zero_words_kernel", st, (double)(a_bytes + b_bytes), kWorkBytes);
    launch_pdl(zero_words_kernel, dim3(1), dim3(128), 0, st, static_cast<uint32_t*>(a), (int)(a ? a_bytes / 4 : 0), static_cast<uint32_t*>(b),
               (int)(b ? b_bytes / 4 : 0));
    return ls.done();
}

}  // namespace fi
