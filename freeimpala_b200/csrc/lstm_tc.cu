// FarmerLstm recurrence (reference cmd/libtorch_bench/main.cpp:25-27: nn::LSTM(162 -> 128), batch_first) on the tcgen05 tensor
// cores: the h_{s-1} W_hh^T product of the forward recurrence and the dG_s W_hh product of BPTT, fp32-accurate through the
// 3xFP16 operand pairs of gemm_tc.cu (x * scale = hi + lo' / 2048; main = A_hi B_hi, corr = A_hi B_lo + A_lo B_hi).
//
// The recurrence is a chain of T dependent [rows,128] x [128,512] products: rows are the only free parallelism, and a UMMA
// tile wants 128 of them. One CLUSTER of 8 CTAs owns 128 batch rows for all T steps; CTA c of the cluster owns hidden units
// [16c, 16c+16) = 64 gate columns (i, f, g, o of its units) and keeps its slice of W_hh in shared memory as fp16 pairs for
// the whole kernel. What the CTAs exchange each step goes through distributed shared memory, not through L2:
//
//   forward   every CTA needs all of h_{s-1} as its A operand. The epilogue threads of CTA c write their 16 columns of h_s
//             (fp16 hi/lo', already in the 128-byte-swizzled K-major layout the MMA reads) into the A tiles of all 8 CTAs
//             (st.shared::cluster), then one cluster barrier; A is double-buffered so one barrier per step is enough.
//   backward  dh_{s-1} = dG_s W_hh sums over all 512 gate columns. CTA c multiplies ITS 64 columns of dG_s (its own
//             units: no exchange needed on the input side) with its 64 rows of W_hh into a partial [128,128], and sends
//             each CTA d the 16 output columns d owns (fp32, st.shared::cluster); after the cluster barrier every CTA sums
//             its 8 partials. The partial buffers are double-buffered for the same reason.
//
// Thread roles: 16 epilogue warps (warp w: TMEM lane quarter w & 3 = rows 32 (w & 3) .. +31, unit group w >> 2 = 4 of the
// CTA's 16 units) + 1 warp that allocates TMEM and issues the MMAs. Per step and CTA the tensor core works 768 clocks; the
// rest of the step is the gate math of the epilogue threads and the exchange.
//
// h is bounded by 1, so its fp16 scale is the constant 2^13 and the forward kernel writes h_{s-1} directly as the fp16 pairs
// the W_hh weight-gradient product reads (no max|x| pass, no split pass, no fp32 copy). The scale of a dG row is derived per
// (row, CTA) from that row's own 64 values, so rows with tiny gradients keep their relative accuracy.
#include <cuda_fp16.h>

#include <cuda.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "learner.cuh"
#include "tc_ptx.cuh"

namespace fi {

namespace {

constexpr int kG4 = 4 * kLstmH;                 // 512 gate columns
constexpr int kLtRows = 128;                    // batch rows per cluster = UMMA M
constexpr int kLtCtas = 8;                      // CTAs per cluster
constexpr int kLtUnits = kLstmH / kLtCtas;      // 16 hidden units per CTA
constexpr int kLtCols = 4 * kLtUnits;           // 64 gate columns per CTA
constexpr int kLtEpiWarps = 16;
constexpr int kLtEpiThreads = 32 * kLtEpiWarps;
constexpr int kLtThreads = kLtEpiThreads + 32;  // + the MMA warp
constexpr float kLtHScale = 8192.f;             // hscale_from_bound(1): |h| <= 1
constexpr float kLoInv = 1.f / 2048.f;

// ---- shared-memory maps (byte offsets from a 1024-byte aligned base; identical in every CTA of the cluster) ----
// forward: A[part][kb][half] = h tiles, MN-major [64 units][64 batch rows] fp16 (128-byte swizzle), part 0 = hi, 1 = lo';
// B[kb] = [64 rows W_hi | 64 rows W_lo'][64 k] (K-major, 128-byte swizzle); stage[buf][part] = this CTA's 16 units of h_s in
// the A layout (exchange); ring[buf] = 4 gate tiles; cst[buf]; hrow[buf][part] = the same h row-major (for the global array)
constexpr uint32_t kTile = 128 * 128;                       // one [128 rows][128 bytes] tile
constexpr uint32_t kSlice = 128 * 32;                       // [128 rows][16 units] fp16 = this CTA's units of h (one part)
constexpr uint32_t kGateTile = 128 * 64;                    // [128 rows][16 units] fp32 = one gate (or c) of this CTA's units
constexpr uint32_t kRingBuf = 4 * kGateTile;                // x-projection in, gates out: 4 gate tiles
constexpr uint32_t kFwdABytes = 4 * kTile;                  // A: [part][k-block][batch-row half][64 k-rows][128 B]
constexpr uint32_t kFwdA = 0, kFwdB = kFwdABytes, kFwdStage = kFwdB + 2 * kTile, kFwdRing = kFwdStage + 4 * kSlice;
constexpr uint32_t kFwdCst = kFwdRing + 2 * kRingBuf, kFwdHrow = kFwdCst + 2 * kGateTile, kFwdMisc = kFwdHrow + 4 * kSlice;
constexpr size_t kFwdSmem = kFwdMisc + 256 + 1024;
// backward: A[part] = dG tiles [128 rows][64 j] fp16; B = [128 rows W_hi | 128 rows W_lo'][64 j] (rows = output units);
// red[buf][src][unit group][row][4] fp32 partials
constexpr uint32_t kBwdA = 0, kBwdB = 2 * kTile, kBwdRed = 4 * kTile, kBwdRedBuf = kLtCtas * kLtRows * kLtUnits * 4;
constexpr uint32_t kBwdRow = kBwdRed + 2 * kBwdRedBuf, kBwdMisc = kBwdRow + 4 * kLtRows * 4;
constexpr size_t kBwdSmem = kBwdMisc + 256 + 1024;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void st_cluster_v4f(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// x * scale as an fp16 pair: hi = rn(x s), lo' = rn((x s - hi) * 2048)
__device__ __forceinline__ void split_h(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn((v - __half2float(hi)) * 2048.f);
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
// byte offset of the 16-byte chunk `chunk` (8 fp16 of k) of row `r` inside a 128-byte-swizzled [rows][128 B] tile
__device__ __forceinline__ uint32_t swz(int r, int chunk) { return (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4); }

// max |w| over this CTA's 64 rows of W_hh (gate g, unit 16 rank + u), all 128 columns -> the CTA's weight scale
__device__ float cta_weight_scale(const float* __restrict__ whh, int rank, float* scratch) {
    float mx = 0.f;
    for (int i = threadIdx.x; i < kLtCols * (kLstmH / 4); i += kLtThreads) {
        const int n = i / (kLstmH / 4), c4 = i % (kLstmH / 4);
        const int row = (n >> 4) * kLstmH + rank * kLtUnits + (n & 15);
        const float4 v = __ldg(reinterpret_cast<const float4*>(whh + (size_t)row * kLstmH) + c4);
        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    mx = __uint_as_float(__reduce_max_sync(0xFFFFFFFFu, __float_as_uint(mx)));
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = mx;
    __syncthreads();
    float all = 0.f;
    for (int w = 0; w < kLtThreads / 32; w++) all = fmaxf(all, scratch[w]);
    __syncthreads();
    return hscale_from_bound(all);
}

// FI_LSTM_TRACE=1: clock64 at the phase boundaries of every step, CTA 0, first epilogue thread (points 0..7) and MMA warp
// (points 8..11); the host prints the median phase lengths after each launch (diagnostics; forces a stream sync).
constexpr int kTracePoints = 12, kTraceSteps = 128;
__device__ __forceinline__ void trace_ev(unsigned long long* tr, int s, int point) {
#if FI_TRACE_BUILD
    if (tr && s < kTraceSteps) tr[s * kTracePoints + point] = clock64();
#else
    (void)tr; (void)s; (void)point;
#endif
}

constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B (K-major tiles)
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t saddr) { return ((uint64_t)kDescHi << 32) | ((saddr >> 4) & 0x3FFFu); }
// MN-major fp16 tiles, SWIZZLE_128B: boxes of [64 k-rows][64 mn] = 8 KB (LBO between the two 64-row halves of M), 8-k-row groups
// 1024 B apart (SBO); a k-slice of 16 = 2048 B. (A K-major operand of 32-byte rows -- SWIZZLE_32B, which would make a CTA's 16
// units one contiguous tile as well -- was measured first: the MMAs ran ~5x slower on it.)
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t saddr) {
    return ((uint64_t)kDescHi << 32) | ((8192u >> 4) << 16) | ((saddr >> 4) & 0x3FFFu);
}
// 1-D bulk copies between global and shared memory (sizes and addresses multiples of 16 bytes)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1),
                 "r"(c2)
                 : "memory");
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// Forward. gates[(b,s), 512] holds W_ih z + b_ih on entry and the post-activation gates i,f,g,o on exit (as the fp32 kernel
// of model_farmer.cu). hp_hi / hp_lo [(b,s), 128]: h_{s-1} as fp16 pairs at scale 2^13 (zeros at s = 0); cst [(b,s), 128]: c_s;
// feat[b, 0..127] = h_{T-1}.
//
// Nothing on the per-step path goes through the threads' own global or remote accesses (a TMEM lane is a batch row, so every
// such access would touch 32 different lines per warp instruction: measured 7500 clocks per step for the stores alone):
//   * global memory <-> shared memory by TMA over 3-D views (columns, step, batch row) of the arrays: the x-projection of step
//     s+2 is loaded while step s runs; gates, c and h leave from staging tiles the threads fill with conflict-free stores;
//   * h between the CTAs: the A operand is MN-major (unit-major: row k of a tile holds unit k for 64 batch rows), so the 16
//     units CTA c produces are two contiguous 2 KB pieces per part (hi, lo'). The threads write them into a staging tile;
//     32 bulk copies (cp.async.bulk.shared::cluster) move them into every CTA's A tile and credit the bytes to the
//     destination's mbarrier. A destination's tile is free once its MMA of the step has completed: every CTA tells all CTAs
//     with a remote mbarrier arrive.
// Warp 16 issues the MMAs and all TMA traffic; warps 0..15 do the gate math.
__global__ void __cluster_dims__(kLtCtas, 1, 1) __launch_bounds__(kLtThreads, 1)
lstm_forward_tc_kernel(float* __restrict__ gates, float* __restrict__ cst, int nblk, const __grid_constant__ CUtensorMap map_hp_hi,
                       const __grid_constant__ CUtensorMap map_hp_lo, const float* __restrict__ whh,
                       const float* __restrict__ b_hh, int m, int t, __half* __restrict__ hp_hi,
                       __half* __restrict__ hp_lo, HScale* __restrict__ hp_hs, float* __restrict__ feat, int ldfeat, int live,
                       unsigned long long* trace) {
    extern __shared__ uint8_t lstm_tc_smem_raw[];
    const uint32_t smem = (smem_u32(lstm_tc_smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_g = lstm_tc_smem_raw + (smem - smem_u32(lstm_tc_smem_raw));
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
    const int rank = (int)cluster_ctarank();
    const int row0 = (blockIdx.x / kLtCtas) * live;   // live = 64 or 128 batch rows per cluster (rows >= live of the MMA tile idle)
    const int nh = live >> 6;
    const uint32_t mma_bar = smem + kFwdMisc, tmem_slot = smem + kFwdMisc + 8;
    const uint32_t h_full = smem + kFwdMisc + 16;    // all 8 CTAs' pieces of h_{s-1} have landed in A (64 KB of transactions)
    const uint32_t a_free = smem + kFwdMisc + 24;    // all 8 CTAs' MMAs of the step have completed: their A tiles may be overwritten
    const uint32_t gx_full0 = smem + kFwdMisc + 32;  // [2]: the x-projection of a step has landed in ring buffer b
    float* scratch = reinterpret_cast<float*>(smem_g + kFwdMisc + 64);

    const float w_scale = cta_weight_scale(whh, rank, scratch);
    // B tiles: local gate column n = 16 ug + 4 gate + u4 <-> W_hh row gate * 128 + 16 rank + 4 ug + u4
    for (int i = tid; i < kLtCols * 16; i += kLtThreads) {
        const int n = i >> 4, ch = i & 15;   // 8 k per chunk
        const int row = ((n >> 2) & 3) * kLstmH + rank * kLtUnits + (n >> 4) * 4 + (n & 3);
        const float4 a = __ldg(reinterpret_cast<const float4*>(whh + (size_t)row * kLstmH + ch * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(whh + (size_t)row * kLstmH + ch * 8 + 4));
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        __half hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; j++) split_h(v[j] * w_scale, hi[j], lo[j]);
        const uint32_t off = kFwdB + (uint32_t)(ch >> 3) * kTile + swz(n, ch & 7);
        *reinterpret_cast<uint4*>(smem_g + off) = make_uint4(pack_h2(hi[0], hi[1]), pack_h2(hi[2], hi[3]), pack_h2(hi[4], hi[5]), pack_h2(hi[6], hi[7]));
        *reinterpret_cast<uint4*>(smem_g + off + kLtCols * 128) =
            make_uint4(pack_h2(lo[0], lo[1]), pack_h2(lo[2], lo[3]), pack_h2(lo[4], lo[5]), pack_h2(lo[6], lo[7]));
    }
    // h_{-1} = 0: the whole A tile (hi and lo')
    for (int i = tid; i < (int)(kFwdABytes / 16); i += kLtThreads) reinterpret_cast<uint4*>(smem_g + kFwdA)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(mma_bar, 1);
        mbar_init(h_full, 1);
        mbar_init(a_free, kLtCtas);
        mbar_init(gx_full0, 1);
        mbar_init(gx_full0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (blockIdx.x == 0 && hp_hs) {
            hp_hs->scale = kLtHScale;
            hp_hs->inv = 1.f / kLtHScale;
            hp_hs->amax = 1.f;
            hp_hs->bound = 1.f;
        }
    }
    if (warp == kLtEpiWarps) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_proxy();
    tc_fence_before();
    __syncthreads();
    cluster_arrive();   // every CTA's tiles and barriers are initialised before any peer's copy or arrive can reach them
    cluster_wait();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
    const int col0 = rank * kLtUnits;   // this CTA's first hidden unit = its column inside each gate block

    if (warp == kLtEpiWarps) {
        // ===================== MMA + TMA warp =====================
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
        constexpr uint32_t idesc_wide = umma_idesc(128, 2 * kLtCols, 1, 0, 1);   // h_hi [W_hi | W_lo'] -> [main | corr]; A MN-major
        constexpr uint32_t idesc_half = umma_idesc(128, kLtCols, 1, 0, 1);       // h_lo' W_hi -> corr
        const bool leader = elect_one();
        unsigned long long* tr = (blockIdx.x == 0 && lane == 0) ? trace : nullptr;
        // this CTA's blocks of the blocked arrays (fi_internal.cuh): two 64-row blocks per step, the second one may not exist
        const int blk0 = row0 >> 6, nhalf = min(nh, nblk - blk0);
        auto gate_block = [&](int s, int h) { return gates + (((size_t)s * nblk + blk0 + h) * 8 + rank) * kStepGateBlock; };
        auto cell_block = [&](int s, int h) { return cst + (((size_t)s * nblk + blk0 + h) * 8 + rank) * kStepCellBlock; };
        auto load_gx = [&](int s) {   // x-projection of step s -> ring buffer s & 1: one 16 KB bulk copy per 64-row block
            const uint32_t bar = gx_full0 + (uint32_t)(s & 1) * 8u, dst = smem + kFwdRing + (uint32_t)(s & 1) * kRingBuf;
            mbar_expect_tx(bar, (uint32_t)nhalf * (kRingBuf / 2));
            for (int h = 0; h < nhalf; h++) bulk_load(dst + (uint32_t)h * (kRingBuf / 2), gate_block(s, h), kRingBuf / 2, bar);
        };
        if (leader) {
            load_gx(0);
            if (t > 1) load_gx(1);
        }
        __syncwarp();
        for (int s = 0; s < t; s++) {
            trace_ev(tr, s, 8);
            if (s > 0) {   // h_{s-1}: 32 bulk copies of 2 KB
                if (leader) mbar_expect_tx(h_full, (uint32_t)nh * (kFwdABytes / 2));
                __syncwarp();
                mbar_wait(h_full, (uint32_t)((s - 1) & 1));
            }
            trace_ev(tr, s, 9);
            tc_fence_after();
            const uint32_t a_hi = smem + kFwdA, a_lo = a_hi + kFwdABytes / 2, b = smem + kFwdB;
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 8; ks++) {   // k-slice ks = the 16 units of CTA ks: 16 k-rows of the MN-major tiles
                    const uint32_t ao = (uint32_t)(ks >> 2) * kTile + (uint32_t)(ks & 3) * 2048u;
                    const uint64_t db = kmajor_desc(b + (uint32_t)(ks >> 2) * kTile + (uint32_t)(ks & 3) * 32u);
                    tc_mma_f16(tmem_u, mnmajor_desc(a_hi + ao), db, idesc_wide, ks ? 1u : 0u);
                    tc_mma_f16(tmem_u + kLtCols, mnmajor_desc(a_lo + ao), db, idesc_half, 1u);
                }
                tc_commit(mma_bar);
            }
            __syncwarp();
            trace_ev(tr, s, 10);
            named_bar_sync(1, kLtThreads);   // the gate warps have filled this step's staging tiles
            trace_ev(tr, s, 11);
            if (leader) {
                const uint32_t ring = smem + kFwdRing + (uint32_t)(s & 1) * kRingBuf, cell = smem + kFwdCst + (uint32_t)(s & 1) * kGateTile;
                for (int h = 0; h < nhalf; h++) {
                    bulk_store(gate_block(s, h), ring + (uint32_t)h * (kRingBuf / 2), kRingBuf / 2);
                    bulk_store(cell_block(s, h), cell + (uint32_t)h * (kGateTile / 2), kGateTile / 2);
                }
                if (s + 1 < t) {
                    tma_store_3d(&map_hp_hi, smem + kFwdHrow + (uint32_t)(s & 1) * 2 * kSlice, col0, s + 1, row0);
                    tma_store_3d(&map_hp_lo, smem + kFwdHrow + (uint32_t)(s & 1) * 2 * kSlice + kSlice, col0, s + 1, row0);
                }
                tma_store_commit();
                if (s + 2 < t) {
                    tma_store_wait_read();   // the staging tiles of this step have been read: the ring buffer can take step s + 2
                    load_gx(s + 2);
                }
            }
            __syncwarp();
        }
        if (leader) tma_store_wait_all();
        __syncwarp();
    } else {
        // ===================== gate math (16 warps) =====================
        const int q = warp & 3, ug = warp >> 2;
        const int r = q * 32 + lane, b = row0 + r;
        const bool active = q * 32 < live;   // warp-uniform: with 64 live rows the warps of TMEM lanes 64..127 only keep the barriers
        const bool valid = active && b < m;
        const int unit0 = col0 + ug * 4;   // first of this thread's 4 hidden units
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ug * 16);
        const float inv = 1.f / (kLtHScale * w_scale);
        if (!active) {
            for (int s = 0; s < t; s++) named_bar_sync(1, kLtThreads);
        } else {
        float4 bias[4];
#pragma unroll
        for (int g = 0; g < 4; g++) bias[g] = __ldg(reinterpret_cast<const float4*>(b_hh + g * kLstmH + unit0));
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        if (valid) {   // h_{-1} = 0
            *reinterpret_cast<uint2*>(hp_hi + (size_t)b * t * kLstmH + unit0) = make_uint2(0u, 0u);
            *reinterpret_cast<uint2*>(hp_lo + (size_t)b * t * kLstmH + unit0) = make_uint2(0u, 0u);
        }
        // this thread's 16 bytes of a [128 rows][16 units] fp32 tile (64-byte swizzle) and its 8 bytes of a [128 rows][16 units]
        // fp16 tile (32-byte swizzle): the layouts the TMA boxes have in shared memory; no bank conflicts across a warp's rows
        // (gate tiles: [64-row half][gate][64 rows][64 B], the layout of the global blocks; c tiles: [half][64 rows][64 B])
        const uint32_t row64 = (uint32_t)(r & 63) * 64u + (uint32_t)((ug ^ ((r >> 1) & 3)) << 4);
        const uint32_t offg = (uint32_t)(r >> 6) * (kRingBuf / 2) + row64, offc = (uint32_t)(r >> 6) * (kGateTile / 2) + row64;
        const uint32_t off32 = (uint32_t)r * 32u + (uint32_t)((((ug >> 1) ^ (r >> 2)) & 1) << 4) + (uint32_t)(ug & 1) * 8u;
        // exchange staging, MN-major: unit u is k-row u of a [16 k-rows][64 batch rows] piece; 16-byte chunks swizzled by the k-row
        uint32_t offmn[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int u = ug * 4 + j, rr = r & 63;
            offmn[j] = (uint32_t)(r >> 6) * 2048u + (uint32_t)u * 128u + (uint32_t)((((rr >> 3) ^ u) & 7) << 4) + (uint32_t)(rr & 7) * 2u;
        }
        unsigned long long* tr = (blockIdx.x == 0 && tid == 0) ? trace : nullptr;
        for (int s = 0; s < t; s++) {
            trace_ev(tr, s, 0);
            mbar_wait(mma_bar, (uint32_t)(s & 1));
            trace_ev(tr, s, 1);
            tc_fence_after();
            if (warp == 0 && lane < kLtCtas && s + 1 < t) mbar_arrive_cluster(map_to_cta(a_free, (uint32_t)lane));   // our A tile is free
            uint32_t mn[16], cr[16];
            tmem_ld16(taddr, mn);
            tmem_ld16(taddr + kLtCols, cr);
            tmem_ld_wait();
            tc_fence_before();
            mbar_wait(gx_full0 + (uint32_t)(s & 1) * 8u, (uint32_t)((s >> 1) & 1));
            uint8_t* ring = smem_g + kFwdRing + (uint32_t)(s & 1) * kRingBuf;
            trace_ev(tr, s, 2);
            float pre[4][4];
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const float4 gx = *reinterpret_cast<const float4*>(ring + g * (kGateTile / 2) + offg);
                const float gxa[4] = {gx.x, gx.y, gx.z, gx.w}, ba[4] = {bias[g].x, bias[g].y, bias[g].z, bias[g].w};
#pragma unroll
                for (int u = 0; u < 4; u++)
                    pre[g][u] = fmaf(fmaf(__uint_as_float(cr[g * 4 + u]), kLoInv, __uint_as_float(mn[g * 4 + u])), inv, ba[u]) + gxa[u];
            }
            float ig[4], fg[4], gg[4], og[4], h[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                ig[u] = sigmoidf_(pre[0][u]);
                fg[u] = sigmoidf_(pre[1][u]);
                gg[u] = tanhf(pre[2][u]);
                og[u] = sigmoidf_(pre[3][u]);
                c[u] = fmaf(fg[u], c[u], ig[u] * gg[u]);
                h[u] = og[u] * tanhf(c[u]);
            }
            __half hh[4], hl[4];
#pragma unroll
            for (int u = 0; u < 4; u++) split_h(h[u] * kLtHScale, hh[u], hl[u]);
            trace_ev(tr, s, 3);
            // staging: the gates in place of the x-projection, c, h (row-major for the global array, unit-major for the exchange)
            *reinterpret_cast<float4*>(ring + offg) = make_float4(ig[0], ig[1], ig[2], ig[3]);
            *reinterpret_cast<float4*>(ring + (kGateTile / 2) + offg) = make_float4(fg[0], fg[1], fg[2], fg[3]);
            *reinterpret_cast<float4*>(ring + 2 * (kGateTile / 2) + offg) = make_float4(gg[0], gg[1], gg[2], gg[3]);
            *reinterpret_cast<float4*>(ring + 3 * (kGateTile / 2) + offg) = make_float4(og[0], og[1], og[2], og[3]);
            *reinterpret_cast<float4*>(smem_g + kFwdCst + (uint32_t)(s & 1) * kGateTile + offc) = make_float4(c[0], c[1], c[2], c[3]);
            if (s + 1 < t) {
                uint8_t* hrow = smem_g + kFwdHrow + (uint32_t)(s & 1) * 2 * kSlice;
                *reinterpret_cast<uint2*>(hrow + off32) = make_uint2(pack_h2(hh[0], hh[1]), pack_h2(hh[2], hh[3]));
                *reinterpret_cast<uint2*>(hrow + kSlice + off32) = make_uint2(pack_h2(hl[0], hl[1]), pack_h2(hl[2], hl[3]));
                uint8_t* stage = smem_g + kFwdStage + (uint32_t)(s & 1) * 2 * kSlice;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    *reinterpret_cast<__half*>(stage + offmn[j]) = hh[j];
                    *reinterpret_cast<__half*>(stage + kSlice + offmn[j]) = hl[j];
                }
            } else if (valid) {
                *reinterpret_cast<float4*>(feat + (size_t)b * ldfeat + unit0) = make_float4(h[0], h[1], h[2], h[3]);
            }
            fence_async_proxy();   // TMA stores and bulk copies read the staging tiles through the async proxy
            named_bar_sync(1, kLtThreads);
            trace_ev(tr, s, 4);
            if (s + 1 < t) {
                mbar_wait(a_free, (uint32_t)(s & 1));   // every CTA's MMA of step s has completed
                trace_ev(tr, s, 5);
                // 16 copies per 64 live rows, spread over the active warps: copy index = 2 * destination + part, lane = 64-row half
                const int wi = (warp >> 2) * (live >> 5) + q, per = 16 / ((live >> 5) * 4);   // wi-th active warp; copies per warp
                for (int ci = wi * per; ci < (wi + 1) * per; ci++)
                if (lane < nh) {
                    const uint32_t d = (uint32_t)ci >> 1, part = (uint32_t)ci & 1u, piece = (uint32_t)lane;
                    const uint32_t src = smem + kFwdStage + (uint32_t)(s & 1) * 2 * kSlice + part * kSlice + piece * 2048u;
                    const uint32_t dst = map_to_cta(smem + kFwdA + part * (kFwdABytes / 2) + (uint32_t)(rank >> 2) * kTile + piece * 8192u +
                                                        (uint32_t)(rank & 3) * 2048u, d);
                    const uint32_t bar = map_to_cta(h_full, d);
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                                 "r"(src), "r"(2048u), "r"(bar)
                                 : "memory");
                }
                __syncwarp();
            }
            trace_ev(tr, s, 6);
            trace_ev(tr, s, 7);
        }
        }   // active
    }
    // no CTA may exit while a peer's copy or arrive can still address its shared memory (every copy sent has been waited for
    // by its destination's MMA warp before that CTA gets here)
    tc_fence_before();
    cluster_arrive();
    cluster_wait();
    if (warp == kLtEpiWarps) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}
// ---------------------------------------------------------------------------------------------------------------------
// BPTT. On entry gates holds the post-activation i,f,g,o; on exit the pre-activation gate gradients dG. dfeat [m, ldf]: its
// first 128 columns are dL/dh_{T-1}. dg_hs->amax receives max |dG| (atomicMax; zero it before) for the split that feeds the
// two weight-gradient products.
__global__ void __cluster_dims__(kLtCtas, 1, 1) __launch_bounds__(kLtThreads, 1)
lstm_backward_tc_kernel(float* __restrict__ gates, const float* __restrict__ whh, const float* __restrict__ cst, int nblk,
                        const float* __restrict__ dfeat, int ldf, int m, int t, HScale* __restrict__ dg_hs, float* __restrict__ bias_part,
                        unsigned long long* trace) {
    extern __shared__ uint8_t lstm_tc_smem_raw[];
    const uint32_t smem = (smem_u32(lstm_tc_smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_g = lstm_tc_smem_raw + (smem - smem_u32(lstm_tc_smem_raw));
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
    const int rank = (int)cluster_ctarank();
    const int row0 = (blockIdx.x / kLtCtas) * kLtRows;
    const uint32_t mma_bar = smem + kBwdMisc, tmem_slot = smem + kBwdMisc + 8;
    float* scratch = reinterpret_cast<float*>(smem_g + kBwdMisc + 64);
    float* row_amax = reinterpret_cast<float*>(smem_g + kBwdRow);   // [4 unit groups][128 rows]

    const float w_scale = cta_weight_scale(whh, rank, scratch);
    // B tile: row n = 32 ug' + 4 d + u4' <-> output unit 16 d + 4 ug' + u4' (so that a thread's 32 TMEM columns are its 4 units
    // for each of the 8 destination CTAs); reduction index jl = 16 ug + 4 gate + u4 <-> W_hh row gate * 128 + 16 rank + 4 ug + u4
    for (int i = tid; i < kLstmH * 8; i += kLtThreads) {
        const int n = i >> 3, ch = i & 7;
        const int kout = ((n >> 2) & 7) * kLtUnits + (n >> 5) * 4 + (n & 3);
        __half hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int jl = ch * 8 + j;
            const int row = ((jl >> 2) & 3) * kLstmH + rank * kLtUnits + (jl >> 4) * 4 + (jl & 3);
            split_h(__ldg(whh + (size_t)row * kLstmH + kout) * w_scale, hi[j], lo[j]);
        }
        const uint32_t off = kBwdB + swz(n, ch);
        *reinterpret_cast<uint4*>(smem_g + off) = make_uint4(pack_h2(hi[0], hi[1]), pack_h2(hi[2], hi[3]), pack_h2(hi[4], hi[5]), pack_h2(hi[6], hi[7]));
        *reinterpret_cast<uint4*>(smem_g + off + kTile) =
            make_uint4(pack_h2(lo[0], lo[1]), pack_h2(lo[2], lo[3]), pack_h2(lo[4], lo[5]), pack_h2(lo[6], lo[7]));
    }
    if (tid == 0) {
        mbar_init(mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kLtEpiWarps) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_proxy();
    tc_fence_before();
    __syncthreads();
    cluster_arrive();
    cluster_wait();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    if (warp == kLtEpiWarps) {
        // ===================== MMA warp =====================
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
        constexpr uint32_t idesc_wide = umma_idesc(128, 2 * kLstmH, 0, 0, 1);   // dG_hi [W_hi | W_lo'] -> [main | corr]
        constexpr uint32_t idesc_half = umma_idesc(128, kLstmH, 0, 0, 1);       // dG_lo' W_hi -> corr
        const uint32_t a_hi = smem + kBwdA, a_lo = a_hi + kTile, b = smem + kBwdB;
        unsigned long long* tr = (blockIdx.x == 0 && lane == 0) ? trace : nullptr;
        for (int s = t - 1; s > 0; s--) {
            trace_ev(tr, t - 1 - s, 8);
            named_bar_sync(2, kLtThreads);   // the epilogue threads have written dG_s into the A tiles
            trace_ev(tr, t - 1 - s, 9);
            fence_async_proxy();
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                    const uint32_t o = (uint32_t)ks * 32u;
                    tc_mma_f16(tmem_u, kmajor_desc(a_hi + o), kmajor_desc(b + o), idesc_wide, ks ? 1u : 0u);
                    tc_mma_f16(tmem_u + kLstmH, kmajor_desc(a_lo + o), kmajor_desc(b + o), idesc_half, 1u);
                }
                tc_commit(mma_bar);
            }
            __syncwarp();
            trace_ev(tr, t - 1 - s, 10);
            cluster_arrive();
            cluster_wait();
            trace_ev(tr, t - 1 - s, 11);
        }
    } else {
        // ===================== gate-gradient math + exchange (16 warps) =====================
        const int q = warp & 3, ug = warp >> 2;
        const int r = q * 32 + lane, b = row0 + r;
        const bool valid = b < m;
        const int unit0 = rank * kLtUnits + ug * 4;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ug * 32);
        // blocked arrays (fi_internal.cuh): this thread's 4 units of gate g at step s / of c at step s
        auto gate_ptr = [&](int s, int g) { return gates + step_block_offset(s, b, g * kLstmH + unit0, nblk); };
        auto cell_ptr = [&](int s) { return cst + step_cell_offset(s, b, unit0, nblk); };
        float bsum[4][4] = {};   // bias gradient: this thread's row's dG summed over the steps
        float dc[4] = {0.f, 0.f, 0.f, 0.f}, dh[4] = {0.f, 0.f, 0.f, 0.f};
        if (valid) {
            const float4 v = *reinterpret_cast<const float4*>(dfeat + (size_t)b * ldf + unit0);
            dh[0] = v.x; dh[1] = v.y; dh[2] = v.z; dh[3] = v.w;
        }
        float4 gt[4], cs, cp;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        auto load_step = [&](int s) {
#pragma unroll
            for (int g = 0; g < 4; g++)
                gt[g] = valid ? *reinterpret_cast<const float4*>(gate_ptr(s, g)) : zero4;
            cs = valid ? *reinterpret_cast<const float4*>(cell_ptr(s)) : zero4;
            cp = (valid && s > 0) ? *reinterpret_cast<const float4*>(cell_ptr(s - 1)) : zero4;
        };
        load_step(t - 1);
        float run_max = 0.f;
        uint32_t parity = 0;
        // this thread's 16 dG values are reduction indices jl = 16 ug .. 16 ug + 15 = chunks 2 ug and 2 ug + 1 of row r
        const uint32_t a_off0 = kBwdA + swz(r, 2 * ug), a_off1 = kBwdA + swz(r, 2 * ug + 1);
        unsigned long long* tr = (blockIdx.x == 0 && tid == 0) ? trace : nullptr;
        for (int s = t - 1; s >= 0; s--) {
            trace_ev(tr, t - 1 - s, 0);
            const float iga[4] = {gt[0].x, gt[0].y, gt[0].z, gt[0].w}, fga[4] = {gt[1].x, gt[1].y, gt[1].z, gt[1].w};
            const float gga[4] = {gt[2].x, gt[2].y, gt[2].z, gt[2].w}, oga[4] = {gt[3].x, gt[3].y, gt[3].z, gt[3].w};
            const float csa[4] = {cs.x, cs.y, cs.z, cs.w}, cpa[4] = {cp.x, cp.y, cp.z, cp.w};
            float d[4][4];   // [gate][unit]
            float mx = 0.f;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float tc = tanhf(csa[u]);
                const float dct = dc[u] + dh[u] * oga[u] * (1.f - tc * tc);
                d[0][u] = dct * gga[u] * iga[u] * (1.f - iga[u]);
                d[1][u] = dct * cpa[u] * fga[u] * (1.f - fga[u]);
                d[2][u] = dct * iga[u] * (1.f - gga[u] * gga[u]);
                d[3][u] = dh[u] * tc * oga[u] * (1.f - oga[u]);
                dc[u] = dct * fga[u];
                mx = fmaxf(fmaxf(mx, fmaxf(fabsf(d[0][u]), fabsf(d[1][u]))), fmaxf(fabsf(d[2][u]), fabsf(d[3][u])));
            }
            run_max = fmaxf(run_max, mx);
            float inv_row = 0.f;
            trace_ev(tr, t - 1 - s, 1);
            if (s > 0) {
                // scale of this row's 64 values in this CTA
                row_amax[ug * kLtRows + r] = mx;
                named_bar_sync(1, kLtEpiThreads);
                const float rmx = fmaxf(fmaxf(row_amax[r], row_amax[kLtRows + r]), fmaxf(row_amax[2 * kLtRows + r], row_amax[3 * kLtRows + r]));
                const float sc = hscale_from_bound(rmx);
                inv_row = 1.f / (sc * w_scale);
                __half hi[16], lo[16];
#pragma unroll
                for (int g = 0; g < 4; g++)
#pragma unroll
                    for (int u = 0; u < 4; u++) split_h(d[g][u] * sc, hi[g * 4 + u], lo[g * 4 + u]);
                *reinterpret_cast<uint4*>(smem_g + a_off0) = make_uint4(pack_h2(hi[0], hi[1]), pack_h2(hi[2], hi[3]), pack_h2(hi[4], hi[5]), pack_h2(hi[6], hi[7]));
                *reinterpret_cast<uint4*>(smem_g + a_off1) = make_uint4(pack_h2(hi[8], hi[9]), pack_h2(hi[10], hi[11]), pack_h2(hi[12], hi[13]), pack_h2(hi[14], hi[15]));
                *reinterpret_cast<uint4*>(smem_g + a_off0 + kTile) = make_uint4(pack_h2(lo[0], lo[1]), pack_h2(lo[2], lo[3]), pack_h2(lo[4], lo[5]), pack_h2(lo[6], lo[7]));
                *reinterpret_cast<uint4*>(smem_g + a_off1 + kTile) = make_uint4(pack_h2(lo[8], lo[9]), pack_h2(lo[10], lo[11]), pack_h2(lo[12], lo[13]), pack_h2(lo[14], lo[15]));
                trace_ev(tr, t - 1 - s, 2);
                fence_async_proxy();
                named_bar_arrive(2, kLtThreads);
                trace_ev(tr, t - 1 - s, 3);
            }
            if (valid) {   // dG in place of the gates (read by the weight-gradient products and the bias column sums)
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    *reinterpret_cast<float4*>(gate_ptr(s, g)) = make_float4(d[g][0], d[g][1], d[g][2], d[g][3]);
#pragma unroll
                    for (int u = 0; u < 4; u++) bsum[g][u] += d[g][u];
                }
            }
            if (s == 0) break;
            load_step(s - 1);
            trace_ev(tr, t - 1 - s, 4);
            mbar_wait(mma_bar, parity);
            trace_ev(tr, t - 1 - s, 5);
            parity ^= 1u;
            tc_fence_after();
            {
                uint32_t mn[32], cr[32];
                tmem_ld32(taddr, mn);
                tmem_ld32(taddr + kLstmH, cr);
                tmem_ld_wait();
                tc_fence_before();
                // red[buf][src = rank][ug][r][4] in CTA d
                const uint32_t dst = smem + kBwdRed + (uint32_t)(s & 1) * kBwdRedBuf + (uint32_t)(((rank * 4 + ug) * kLtRows + r) * 16);
#pragma unroll
                for (int dcta = 0; dcta < kLtCtas; dcta++) {
                    float p[4];
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        p[u] = fmaf(__uint_as_float(cr[dcta * 4 + u]), kLoInv, __uint_as_float(mn[dcta * 4 + u])) * inv_row;
                    st_cluster_v4f(map_to_cta(dst, (uint32_t)dcta), p[0], p[1], p[2], p[3]);
                }
            }
            trace_ev(tr, t - 1 - s, 6);
            cluster_arrive();
            cluster_wait();
            trace_ev(tr, t - 1 - s, 7);
            // dh_{s-1} of this thread's units = sum of the 8 CTAs' partials
            {
                const uint8_t* red = smem_g + kBwdRed + (uint32_t)(s & 1) * kBwdRedBuf + (uint32_t)((ug * kLtRows + r) * 16);
                float4 acc = *reinterpret_cast<const float4*>(red);
#pragma unroll
                for (int src = 1; src < kLtCtas; src++) {
                    const float4 v = *reinterpret_cast<const float4*>(red + (size_t)src * 4 * kLtRows * 16);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
                dh[0] = acc.x; dh[1] = acc.y; dh[2] = acc.z; dh[3] = acc.w;
            }
        }
        if (dg_hs) {
            const uint32_t wm = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(run_max));
            if (lane == 0 && wm) atomicMax(reinterpret_cast<unsigned int*>(&dg_hs->amax), wm);
        }
        // bias gradient of this cluster's rows: fixed-order sums (shuffle tree over a warp's 32 rows, then the 4 row quarters)
        float* bq = row_amax;   // [4 quarters][64 local columns], reusing the row-scale scratch
        named_bar_sync(1, kLtEpiThreads);
#pragma unroll
        for (int g = 0; g < 4; g++)
#pragma unroll
            for (int u = 0; u < 4; u++) {
                float v = bsum[g][u];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
                if (lane == 0) bq[q * kLtCols + g * kLtUnits + ug * 4 + u] = v;
            }
        named_bar_sync(1, kLtEpiThreads);
        if (tid < kLtCols) {
            const int g = tid >> 4, uu = tid & 15;
            bias_part[(size_t)(blockIdx.x / kLtCtas) * kG4 + g * kLstmH + rank * kLtUnits + uu] =
                (bq[tid] + bq[kLtCols + tid]) + (bq[2 * kLtCols + tid] + bq[3 * kLtCols + tid]);
        }
    }
    tc_fence_before();
    cluster_arrive();
    cluster_wait();
    if (warp == kLtEpiWarps) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------
// Off by default: at the benchmark shape (1024 x 100) the tensor-core recurrence only matches the fp32 FFMA kernels of
// model_farmer.cu (profiles/r2_lstm_tc.md: the per-step exchange of h between the 8 CTAs of a cluster is bound by the ~20 B/clk
// an SM can push into distributed shared memory, and only 15 clusters of 8 CTAs are resident at once). FI_LSTM_TC=1 or
// fi_debug_set_lstm_tc(1) select it; the parity tests run both.
static std::atomic<int> g_lstm_tc{-1};
bool lstm_tc_enabled() {
    const int v = g_lstm_tc.load(std::memory_order_relaxed);
    if (v >= 0) return v != 0;
    static const bool on = [] { const char* e = getenv("FI_LSTM_TC"); return e && e[0] == '1'; }();
    return on;
}
void lstm_tc_set(int on) { g_lstm_tc.store(on, std::memory_order_relaxed); }

unsigned long long* lstm_trace_buffer() {
    static unsigned long long* buf = [] {
        unsigned long long* p = nullptr;
        const char* e = getenv("FI_LSTM_TRACE");
        if (e && e[0] == '1' && trace_hooks_built("FI_LSTM_TRACE") && cudaMalloc((void**)&p, sizeof(unsigned long long) * kTracePoints * kTraceSteps) != cudaSuccess) p = nullptr;
        return p;
    }();
    return buf;
}
// median clocks between consecutive trace points (and from the last point of a step to the first of the next)
void lstm_trace_report(const char* name, unsigned long long* dev, int steps, cudaStream_t st) {
    if (!dev || cudaStreamSynchronize(st) != cudaSuccess) return;
    std::vector<unsigned long long> h((size_t)kTracePoints * kTraceSteps);
    if (cudaMemcpy(h.data(), dev, h.size() * sizeof(h[0]), cudaMemcpyDeviceToHost) != cudaSuccess) return;
    steps = steps < kTraceSteps ? steps : kTraceSteps;
    auto median = [](std::vector<long long>& v) { std::sort(v.begin(), v.end()); return v.empty() ? 0ll : v[v.size() / 2]; };
    fprintf(stderr, "[lstm trace] %s, %d steps, median clocks:", name, steps);
    for (int role = 0; role < 2; role++) {
        const int p0 = role ? 8 : 0, p1 = role ? 12 : 8;
        fprintf(stderr, "\n  %s", role ? "mma warp:" : "epilogue:");
        for (int p = p0; p < p1; p++) {
            std::vector<long long> d;
            for (int s = 2; s + 2 < steps; s++) {
                const unsigned long long a = h[(size_t)s * kTracePoints + p];
                const unsigned long long b = p + 1 < p1 ? h[(size_t)s * kTracePoints + p + 1] : h[(size_t)(s + 1) * kTracePoints + p0];
                if (a && b) d.push_back((long long)(b - a));
            }
            fprintf(stderr, " p%d->%d %lld", p, p + 1 < p1 ? p + 1 : p0, median(d));
        }
    }
    fprintf(stderr, "\n");
    cudaMemset(dev, 0, h.size() * sizeof(h[0]));
}

// gemm_tc.cu
int make_step_tensor_map(CUtensorMap* map, const void* base, int elem_bytes, uint64_t cols, uint64_t t, uint64_t rows, uint32_t box_cols,
                         uint32_t box_rows, int swizzle_bytes);

namespace {
// Batch rows per cluster: 64 (half of the MMA tile idle, but twice the CTAs share the gate math and the exchange, which are
// what bound a step) while the clusters still fit the GPU in one wave, else 128. FI_LSTM_LIVE=64|128 forces it.
int lstm_live_rows(int m) {
    const char* e = getenv("FI_LSTM_LIVE");   // read per launch (diagnostics and tests switch it between learners)
    const int forced = e ? atoi(e) : 0;
    if (forced == 64 || forced == 128) return forced;
    return ((m + 63) / 64) * kLtCtas <= kNumSMs ? 64 : 128;
}
struct FwdMaps { CUtensorMap hp_hi, hp_lo; };
struct FwdKey {
    const void *hp_hi, *hp_lo; int m, t;
    bool operator==(const FwdKey& o) const { return hp_hi == o.hp_hi && hp_lo == o.hp_lo && m == o.m && t == o.t; }
};
}  // namespace

int launch_lstm_forward_tc(float* gates, const float* whh, const float* b_hh, int m, int t, void* hp_hi, void* hp_lo, HScale* hp_hs,
                           float* cst, float* feat, int ldfeat, cudaStream_t st) {
    // one set of tensor maps per (arrays, batch rows, steps): built on the first step of a workspace, reused afterwards
    static std::mutex mu;
    static std::vector<std::pair<FwdKey, FwdMaps>> cache;
    const int live_rows = lstm_live_rows(m);
    const FwdKey key{hp_hi, hp_lo, m * 256 + live_rows, t};
    FwdMaps maps;
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = std::find_if(cache.begin(), cache.end(), [&](const std::pair<FwdKey, FwdMaps>& e) { return e.first == key; });
        if (it == cache.end()) {
            FI_TRY(make_step_tensor_map(&maps.hp_hi, hp_hi, 2, kLstmH, (uint64_t)t, (uint64_t)m, kLtUnits, (uint32_t)live_rows, 32));
            FI_TRY(make_step_tensor_map(&maps.hp_lo, hp_lo, 2, kLstmH, (uint64_t)t, (uint64_t)m, kLtUnits, (uint32_t)live_rows, 32));
            if (cache.size() >= 64) cache.clear();
            cache.emplace_back(key, maps);
        } else {
            maps = it->second;
        }
    }
    LaunchScope ls("lstm_forward_tc_kernel", st, 2.0 * kLstmH * kG4 * (double)m * t, kWorkFlops);
    static std::atomic<uint64_t> attr{0};
    FI_TRY(ensure_dynamic_smem(attr, (const void*)lstm_forward_tc_kernel, (int)kFwdSmem));
    const int live = lstm_live_rows(m), clusters = (m + live - 1) / live;
    if (lstm_trace_buffer()) {   // diagnostics: how many clusters of this kernel fit the GPU at once
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(clusters * kLtCtas);
        cfg.blockDim = dim3(kLtThreads);
        cfg.dynamicSmemBytes = kFwdSmem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = kLtCtas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = -1;
        const cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, (const void*)lstm_forward_tc_kernel, &cfg);
        fprintf(stderr, "[lstm trace] forward: %d clusters of %d CTAs launched (live rows %d), max active clusters %d (%s)\n", clusters, kLtCtas,
                live, nc, cudaGetErrorString(e));
    }
    lstm_forward_tc_kernel<<<clusters * kLtCtas, kLtThreads, kFwdSmem, st>>>(gates, cst, (m + kStepBlockRows - 1) / kStepBlockRows, maps.hp_hi, maps.hp_lo, whh, b_hh, m, t,
                                                                              static_cast<__half*>(hp_hi), static_cast<__half*>(hp_lo), hp_hs,
                                                                              feat, ldfeat, live, lstm_trace_buffer());
    const int rc = ls.done();
    lstm_trace_report("forward", lstm_trace_buffer(), t, st);
    return rc;
}

// db_ih = db_hh = the clusters' partial column sums of dG, added in cluster order
__global__ void lstm_bias_grad_kernel(const float* __restrict__ part, int nparts, float* __restrict__ g_bih, float* __restrict__ g_bhh) {
    pdl_wait();
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= kG4) return;
    float acc = 0.f;
    for (int p = 0; p < nparts; p++) acc += part[(size_t)p * kG4 + col];
    g_bih[col] = acc;
    g_bhh[col] = acc;
}

// dG (blocked array) -> the row-major fp16 pair [m * t, 512] the two weight-gradient products read, scale from hs->amax
__global__ void __launch_bounds__(256)
lstm_split_gates_kernel(const float* __restrict__ gates, int m, int t, int nblk, __half2* __restrict__ hi, __half2* __restrict__ lo, HScale* hs) {
    const float scale = hscale_from_bound(hs->amax);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        hs->scale = scale;
        hs->inv = 1.f / scale;
        hs->bound = hs->amax;
    }
    const size_t pairs = (size_t)m * t * (kG4 / 2);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / (kG4 / 2)), col = (int)(i % (kG4 / 2)) * 2;
        const int b = row / t, s = row - b * t;
        const float2 v = *reinterpret_cast<const float2*>(gates + step_block_offset(s, b, col, nblk));
        const float v0 = v.x * scale, v1 = v.y * scale;
        const __half2 h = __floats2half2_rn(v0, v1);
        const float2 hf = __half22float2(h);
        hi[i] = h;
        lo[i] = __floats2half2_rn(fmaf(v0, 2048.f, -2048.f * hf.x), fmaf(v1, 2048.f, -2048.f * hf.y));
    }
}

int launch_lstm_bias_grad(const float* part, int nparts, float* g_bih, float* g_bhh, cudaStream_t st) {
    LaunchScope ls("lstm_bias_grad_kernel", st, 4.0 * kG4 * (nparts + 2), kWorkBytes);
    launch_pdl(lstm_bias_grad_kernel, dim3(2), dim3(256), 0, st, part, nparts, g_bih, g_bhh);
    return ls.done();
}

int launch_lstm_split_gates(const float* gates, int m, int t, void* hi, void* lo, HScale* hs, cudaStream_t st) {
    LaunchScope ls("lstm_split_gates_kernel", st, 8.0 * m * t * kG4, kWorkBytes);
    lstm_split_gates_kernel<<<kNumSMs * 8, 256, 0, st>>>(gates, m, t, (m + kStepBlockRows - 1) / kStepBlockRows, static_cast<__half2*>(hi),
                                                       static_cast<__half2*>(lo), hs);
    return ls.done();
}

int launch_lstm_backward_tc(float* gates, const float* whh, const float* cst, const float* dfeat, int ldf, int m, int t, HScale* dg_hs,
                            float* bias_part, float* g_bih, float* g_bhh, cudaStream_t st) {
    LaunchScope ls("lstm_backward_tc_kernel", st, 2.0 * kLstmH * kG4 * (double)m * t, kWorkFlops);
    static std::atomic<uint64_t> attr{0};
    FI_TRY(ensure_dynamic_smem(attr, (const void*)lstm_backward_tc_kernel, (int)kBwdSmem));
    const int clusters = (m + kLtRows - 1) / kLtRows;
    lstm_backward_tc_kernel<<<clusters * kLtCtas, kLtThreads, kBwdSmem, st>>>(gates, whh, cst, (m + kStepBlockRows - 1) / kStepBlockRows, dfeat, ldf, m,
                                                                               t, dg_hs, bias_part, lstm_trace_buffer());
    FI_TRY(ls.done());
    lstm_trace_report("backward", lstm_trace_buffer(), t, st);
    return launch_lstm_bias_grad(bias_part, clusters, g_bih, g_bhh, st);
}

}  // namespace fi

// Diagnostics / tests: 1 = run the FarmerLstm recurrence on the tensor cores (lstm_tc.cu), 0 = on the fp32 FFMA kernels,
// -1 = as the environment says (FI_LSTM_TC, default 0). Takes effect at the next step.
extern "C" void fi_debug_set_lstm_tc(int on) { fi::lstm_tc_set(on); }
