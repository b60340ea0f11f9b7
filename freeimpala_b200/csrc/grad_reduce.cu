// One pair of launches finishes every gradient of a learner step that the backward GEMMs leave in pieces:
//   * weight gradients: split-K partial slabs [splits][m*n] of the wgrad products, summed in fixed order (deterministic);
//   * bias gradients: column sums of a matrix -- the per-32-row sums the dgrad epilogue left ([4*ceil(rows/128), 512]) or
//     the head gradient itself ([rows, 17]).
// Round 1 ran reduce_splits after each wgrad product and colsum_partial + colsum_final per bias: 18 launches of 4-12 us,
// 5.7 % of the step. Here every producer keeps its own slab / partial buffer until the end of the
// backward pass and two launches read them all: pass 1 reduces the tall matrices to <= 1024 row-block partials, pass 2
// sums slabs and row-block partials into the gradient arena. HBM-bound: 4 B per partial element read.
#include "fi_internal.cuh"

namespace fi {

__global__ void __launch_bounds__(256)
grad_colsum_partial_kernel(const GradSegTable t) {
    pdl_wait();
    // blocks [first_block1[z], first_block1[z+1]) belong to column-sum segment z; block = rows [r0, r1) x 32 columns; 8 row-lanes x
    // 32 column-lanes
    int z = 0;
    while (z + 1 < t.num_colsum && (int)blockIdx.x >= t.first_block1[z + 1]) z++;
    const GradSeg& g = t.seg[t.colsum_index[z]];
    const int b = blockIdx.x - t.first_block1[z], col_blocks = (g.n + 31) / 32;
    const int by = b / col_blocks, bx = b % col_blocks;
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int col = bx * 32 + cx;
    const int r0 = by * g.rows_per_block, r1 = min(g.splits, r0 + g.rows_per_block);
    float s = 0.f;
    if (col < g.n)
        for (int r = r0 + ry; r < r1; r += 8) s += __ldg(g.src + (size_t)r * g.stride + col);
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && col < g.n) {
        float tsum = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) tsum += red[i][cx];
        g.scratch[(size_t)by * g.n + col] = tsum;
    }
}

__global__ void __launch_bounds__(256)
grad_finalize_kernel(const GradSegTable t) {
    pdl_wait();
    // blocks [first_block[i], first_block[i+1]) belong to segment i
    int i = 0;
    while (i + 1 < t.count && (int)blockIdx.x >= t.first_block[i + 1]) i++;
    const GradSeg& g = t.seg[i];
    const int b = blockIdx.x - t.first_block[i];
    if (g.kind == 0) {   // dst[j] = sum_s src[s * stride + j], 4 outputs per thread (16-byte accesses when aligned)
        const size_t j0 = ((size_t)b * 256 + threadIdx.x) * 4;
        if (j0 >= (size_t)g.n) return;
        if (j0 + 3 < (size_t)g.n && ((g.stride | (size_t)(reinterpret_cast<uintptr_t>(g.src) >> 2) | (size_t)(reinterpret_cast<uintptr_t>(g.dst) >> 2)) & 3) == 0) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int k = 0; k < g.splits; k++) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(g.src + (size_t)k * g.stride + j0));
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            *reinterpret_cast<float4*>(g.dst + j0) = s;
        } else {
            for (size_t j = j0; j < j0 + 4 && j < (size_t)g.n; j++) {
                float s = 0.f;
                for (int k = 0; k < g.splits; k++) s += __ldg(g.src + (size_t)k * g.stride + j);
                g.dst[j] = s;
            }
        }
    } else {             // one warp per column: the row-block partials with a fixed shuffle tree
        const int col = b * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
        if (col >= g.n) return;
        float s = 0.f;
        for (int p = lane; p < g.rb; p += 32) s += __ldg(g.scratch + (size_t)p * g.n + col);
        s = warp_sum(s);
        if (lane == 0) g.dst[col] = s;
    }
}

int grad_table_add_slabs(GradSegTable* t, const float* slabs, int splits, size_t stride, int n, float* dst) {
    if (t->count >= kGradSegMax) return set_error(FI_ERR_STATE, "gradient segment table is full");
    GradSeg& g = t->seg[t->count++];
    g = GradSeg{};
    g.src = slabs; g.dst = dst; g.n = n; g.splits = splits; g.stride = stride; g.kind = 0;
    return FI_OK;
}

size_t grad_colsum_scratch_bytes(int rows, int n) {
    int rb = (rows + 127) / 128;
    if (rb > 1024) rb = 1024;
    if (rb < 1) rb = 1;
    return (size_t)rb * n * sizeof(float);
}

int grad_table_add_colsum(GradSegTable* t, const float* x, int ld, int rows, int n, float* dst, float* scratch) {
    if (t->count >= kGradSegMax || t->num_colsum >= kGradSegMax) return set_error(FI_ERR_STATE, "gradient segment table is full");
    GradSeg& g = t->seg[t->count];
    g = GradSeg{};
    g.src = x; g.dst = dst; g.n = n; g.splits = rows; g.stride = (size_t)ld; g.kind = 1;
    int rb = (rows + 127) / 128;
    if (rb > 1024) rb = 1024;
    if (rb < 1) rb = 1;
    g.rb = rb;
    g.rows_per_block = (rows + rb - 1) / rb;
    g.scratch = scratch;
    t->colsum_index[t->num_colsum++] = t->count++;
    return FI_OK;
}

int launch_grad_finalize(GradSegTable* t, cudaStream_t st) {
    if (t->count == 0) return FI_OK;
    double bytes1 = 0, bytes2 = 0;
    int blocks = 0, blocks1 = 0, z = 0;
    for (int i = 0; i < t->count; i++) {
        const GradSeg& g = t->seg[i];
        t->first_block[i] = blocks;
        if (g.kind == 0) {
            blocks += ceil_div(g.n, 1024);
            bytes2 += 4.0 * g.n * (g.splits + 1);
        } else {
            blocks += ceil_div(g.n, 8);
            bytes1 += 4.0 * ((double)g.splits * g.n + (double)g.rb * g.n);
            bytes2 += 4.0 * g.n * (g.rb + 1);
            t->first_block1[z++] = blocks1;
            blocks1 += g.rb * ceil_div(g.n, 32);
        }
    }
    t->first_block[t->count] = blocks;
    t->first_block1[t->num_colsum] = blocks1;
    if (t->num_colsum > 0) {
        LaunchScope l1("grad_colsum_partial_kernel", st, bytes1, kWorkBytes);
        launch_pdl(grad_colsum_partial_kernel, dim3(blocks1), dim3(256), 0, st, *t);
        FI_TRY(l1.done());
    }
    LaunchScope l2("grad_finalize_kernel", st, bytes2, kWorkBytes);
    launch_pdl(grad_finalize_kernel, dim3(blocks), dim3(256), 0, st, *t);
    return l2.done();
}

}  // namespace fi
