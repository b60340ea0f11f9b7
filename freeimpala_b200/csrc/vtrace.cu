// V-trace targets, policy-gradient advantages and the fused policy/value/entropy loss head.
//
// NOT in the reference: BASELINE.json's north_star adds it. Restates
// Espeholt et al. 2018 (arXiv:1802.01561) eq. (1) + Remark 2 (lambda) and section 4.2:
//   rho_s = min(rho_bar, e^{log_rho_s})      c_s = lambda * min(c_bar, e^{log_rho_s})
//   delta_s = rho_s (r_s + g_s V_{s+1} - V_s),            V_T = bootstrap
//   acc_s = delta_s + g_s c_s acc_{s+1}, acc_T = 0;       vs_s = V_s + acc_s
//   pg_adv_s = min(pg_rho_bar, e^{log_rho_s}) (r_s + g_s vs_{s+1} - V_s),  vs_T = bootstrap
// The reverse-time first-order recurrence is associative under (a1,b1)o(a2,b2) =
// (a1 a2, b1 + a1 b2), so each warp scans 32 time steps at once with shuffles and carries one
// scalar across 32-step chunks (last chunk first).
#include "fi_internal.cuh"

namespace fi {

// Reverse inclusive scan of acc_l = b_l + a_l * acc_{l+1} over the 32 lanes of a warp.
// On return (a,b) is the composition of lanes l..31, i.e. acc_l = b + a * carry_in.
__device__ __forceinline__ void warp_reverse_linear_scan(float& a, float& b, int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float a2 = __shfl_down_sync(0xffffffffu, a, off);
        const float b2 = __shfl_down_sync(0xffffffffu, b, off);
        if (lane + off < 32) {
            b = fmaf(a, b2, b);
            a *= a2;
        }
    }
}

// ---------------------------------------------------------------- standalone scan -------
// Arrays are trajectory-major [m, t] fp32. One CTA stages `tr` consecutive trajectories
// (a contiguous span of tr*t floats per array) into shared memory with 16-byte coalesced
// loads, each warp scans one trajectory out of shared memory, and the two result spans are
// written back with 16-byte coalesced stores. Algorithmic traffic: 16 B in + 8 B out per
// transition (+4 B per trajectory for the bootstrap); HBM-bound.
__device__ __forceinline__ void stage_in(float* __restrict__ dst, const float* __restrict__ src, size_t e0,
                                         int count, bool vec_ok) {
    if (vec_ok) {
        const float4* s4 = reinterpret_cast<const float4*>(src + e0);
        float4* d4 = reinterpret_cast<float4*>(dst);
        const int n4 = count >> 2;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) d4[i] = __ldg(s4 + i);
        for (int i = (n4 << 2) + threadIdx.x; i < count; i += blockDim.x) dst[i] = __ldg(src + e0 + i);
    } else {
        for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = __ldg(src + e0 + i);
    }
}
__device__ __forceinline__ void stage_out(float* __restrict__ dst, const float* __restrict__ src, size_t e0,
                                          int count, bool vec_ok) {
    if (vec_ok) {
        float4* d4 = reinterpret_cast<float4*>(dst + e0);
        const float4* s4 = reinterpret_cast<const float4*>(src);
        const int n4 = count >> 2;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) d4[i] = s4[i];
        for (int i = (n4 << 2) + threadIdx.x; i < count; i += blockDim.x) dst[e0 + i] = src[i];
    } else {
        for (int i = threadIdx.x; i < count; i += blockDim.x) dst[e0 + i] = src[i];
    }
}

__global__ void vtrace_scan_kernel(int m, int t, int tr, const float* __restrict__ log_rho,
                                   const float* __restrict__ discount, const float* __restrict__ reward,
                                   const float* __restrict__ value, const float* __restrict__ bootstrap,
                                   float rho_bar, float c_bar, float pg_rho_bar, float lambda_,
                                   float* __restrict__ vs_out, float* __restrict__ adv_out, int vec_ok) {
    extern __shared__ __align__(16) float smem[];
    const int span = tr * t;  // floats per array per tile (padded to a multiple of 4 below)
    const int span_pad = (span + 3) & ~3;
    float* s_rho = smem;                  // reused for vs
    float* s_disc = smem + span_pad;
    float* s_rew = smem + 2 * span_pad;   // reused for pg_adv
    float* s_val = smem + 3 * span_pad;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ntiles = (m + tr - 1) / tr;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b0 = tile * tr;
        const int ntraj = min(tr, m - b0);
        const int count = ntraj * t;
        const size_t e0 = (size_t)b0 * t;
        const bool v_ok = vec_ok && ((e0 & 3) == 0);
        stage_in(s_rho, log_rho, e0, count, v_ok);
        stage_in(s_disc, discount, e0, count, v_ok);
        stage_in(s_rew, reward, e0, count, v_ok);
        stage_in(s_val, value, e0, count, v_ok);
        __syncthreads();
        if (warp < ntraj) {
            const int o = warp * t;
            const float boot = __ldg(bootstrap + b0 + warp);
            float carry = 0.f;          // acc_{s+1} entering the chunk
            float vs_next_chunk = boot; // vs at the first step of the chunk after this one
            for (int c0 = ((t - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
                const int s = c0 + lane;
                const bool valid = s < t;
                float a = 0.f, b = 0.f, v = 0.f, r = 0.f, g = 0.f, is = 0.f;
                if (valid) {
                    is = expf(s_rho[o + s]);
                    g = s_disc[o + s];
                    r = s_rew[o + s];
                    v = s_val[o + s];
                    const float v_next = (s == t - 1) ? boot : s_val[o + s + 1];
                    const float rho = fminf(rho_bar, is);
                    const float cc = lambda_ * fminf(c_bar, is);
                    b = rho * (r + g * v_next - v);  // delta_s
                    a = g * cc;
                }
                warp_reverse_linear_scan(a, b, lane);
                const float acc = fmaf(a, carry, b);
                const float vs = v + acc;
                float vs_next = __shfl_down_sync(0xffffffffu, vs, 1);
                if (lane == 31 || s == t - 1) vs_next = vs_next_chunk;
                if (valid) {
                    s_rho[o + s] = vs;
                    s_rew[o + s] = fminf(pg_rho_bar, is) * (r + g * vs_next - v);
                }
                carry = __shfl_sync(0xffffffffu, acc, 0);
                vs_next_chunk = __shfl_sync(0xffffffffu, vs, 0);
            }
        }
        __syncthreads();
        stage_out(vs_out, s_rho, e0, count, v_ok);
        if (adv_out) stage_out(adv_out, s_rew, e0, count, v_ok);
        __syncthreads();
    }
}

int launch_vtrace_scan(int m, int t, const float* log_rho, const float* discount, const float* reward,
                       const float* value, const float* bootstrap, float rho_bar, float c_bar,
                       float pg_rho_bar, float lambda_, float* vs, float* pg_adv, cudaStream_t stream) {
    if (m <= 0 || t <= 0) return FI_OK;
    if (!log_rho || !discount || !reward || !value || !bootstrap || !vs)
        return set_error(FI_ERR_ARG, "vtrace: null argument");
    // trajectories per CTA: one warp each, at most 8, sized so that >= 2 CTAs fit one SM
    const size_t per_traj = (size_t)16 * t;  // 4 staged arrays x 4 B
    int tr = (int)((96 * 1024) / per_traj);
    if (tr > 8) tr = 8;
    if (tr < 1) tr = 1;
    const size_t smem = 4 * (((size_t)tr * t + 3) & ~(size_t)3) * sizeof(float);
    if (smem > 227 * 1024) return set_error(FI_ERR_ARG, "vtrace: T=%d too long for shared-memory staging", t);
    static bool attr_set = false;
    if (!attr_set) {
        FI_CUDA_OK(cudaFuncSetAttribute(vtrace_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    uintptr_t al = (uintptr_t)log_rho | (uintptr_t)discount | (uintptr_t)reward | (uintptr_t)value | (uintptr_t)vs |
                   (uintptr_t)pg_adv;
    const int vec_ok = (al & 15) == 0;
    const int ntiles = (m + tr - 1) / tr;
    const int cap = kNumSMs * 16;
    const int grid = ntiles < cap ? ntiles : cap;
    // algorithmic traffic: 16 B in + 8 B out per transition, + 4 B per trajectory (bootstrap)
    LaunchScope ls("vtrace_scan_kernel", stream, (pg_adv ? 24.0 : 20.0) * (double)m * t + 4.0 * m, kWorkBytes);
    vtrace_scan_kernel<<<grid, 32 * tr, smem, stream>>>(m, t, tr, log_rho, discount, reward, value, bootstrap,
                                                         rho_bar, c_bar, pg_rho_bar, lambda_, vs, pg_adv, vec_ok);
    return ls.done();
}

// ---------------------------------------------------------------- fused loss head -------
// One warp per trajectory of a gathered batch [m, t, 256 words]. For every transition the
// lane reads the learner head row (16 logits + value), the record's behaviour logits, action,
// reward and discount, and in ONE reverse pass produces: log-softmax, log_rho, the V-trace
// scan, pg advantages, the three losses and d(total)/d(head) -- no [m,t] intermediates go
// through HBM between "scan" and "loss". Losses are accumulated in double (4 atomics/warp).
__global__ void __launch_bounds__(128)
vtrace_loss_head_kernel(const float* __restrict__ batch, int m, int t, const float* __restrict__ head, int ldh,
                        float rho_bar, float c_bar, float pg_rho_bar, float lambda_, float baseline_cost,
                        float entropy_cost, float* __restrict__ dhead, float* __restrict__ vs_out,
                        float* __restrict__ adv_out, double* __restrict__ losses, float* __restrict__ dhead_hi,
                        float* __restrict__ dhead_lo, int ld_split) {
    const int lane = threadIdx.x & 31;
    const int traj = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (traj >= m) return;
    const float* slot = batch + (size_t)traj * t * kRecWords;
    const float boot = __ldg(slot + (size_t)(t - 1) * kRecWords + kWAux);
    float carry = 0.f, vs_next_chunk = boot, v_next_chunk = boot;
    double l_pg = 0.0, l_bl = 0.0, l_ent = 0.0;
    for (int c0 = ((t - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
        const int s = c0 + lane;
        const bool valid = s < t;
        const size_t row = (size_t)traj * t + (valid ? s : t - 1);
        const float* rec = slot + (size_t)(valid ? s : t - 1) * kRecWords;
        const float* h = head + row * ldh;
        float z[kNumActions], p[kNumActions];
        float mx = -INFINITY, mmx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kNumActions; j++) {
            z[j] = __ldg(h + j);
            mx = fmaxf(mx, z[j]);
        }
        const float v = __ldg(h + kNumActions);
        int act = __float_as_int(__ldg(rec + kWAction));
        act = act < 0 ? 0 : (act >= kNumActions ? kNumActions - 1 : act);
        float se = 0.f, mu_se = 0.f, mu_a = 0.f, z_a = 0.f;
#pragma unroll
        for (int j = 0; j < kNumActions; j++) {
            const float mu = __ldg(rec + kWMu + j);
            p[j] = mu;  // parked until the behaviour max is known
            mmx = fmaxf(mmx, mu);
            se += expf(z[j] - mx);
            if (j == act) { mu_a = mu; z_a = z[j]; }
        }
#pragma unroll
        for (int j = 0; j < kNumActions; j++) mu_se += expf(p[j] - mmx);
        const float lse = mx + logf(se);
        const float logp_a = z_a - lse;
        const float log_mu_a = mu_a - (mmx + logf(mu_se));
        float plogp = 0.f;
#pragma unroll
        for (int j = 0; j < kNumActions; j++) {
            const float lp = z[j] - lse;
            p[j] = expf(lp);
            z[j] = lp;  // z now holds log pi
            plogp = fmaf(p[j], lp, plogp);
        }
        const float r = __ldg(rec + kWReward), g = __ldg(rec + kWDiscount);
        const float is = expf(logp_a - log_mu_a);
        float v_next = __shfl_down_sync(0xffffffffu, v, 1);
        if (lane == 31 || s >= t - 1) v_next = v_next_chunk;
        float a = 0.f, b = 0.f;
        if (valid) {
            b = fminf(rho_bar, is) * (r + g * v_next - v);
            a = g * lambda_ * fminf(c_bar, is);
        }
        warp_reverse_linear_scan(a, b, lane);
        const float acc = fmaf(a, carry, b);
        const float vs = v + acc;
        float vs_next = __shfl_down_sync(0xffffffffu, vs, 1);
        if (lane == 31 || s >= t - 1) vs_next = vs_next_chunk;
        const float adv = fminf(pg_rho_bar, is) * (r + g * vs_next - v);
        if (valid) {
            float dv[kHead];
#pragma unroll
            for (int j = 0; j < kNumActions; j++) {
                const float d_pg = adv * (p[j] - (j == act ? 1.f : 0.f));
                const float d_ent = p[j] * (z[j] - plogp);
                dv[j] = fmaf(entropy_cost, d_ent, d_pg);
            }
            dv[kNumActions] = -baseline_cost * (vs - v);
            if (dhead) {
                float* dh = dhead + row * ldh;
#pragma unroll
                for (int j = 0; j < kHead; j++) dh[j] = dv[j];
            }
            if (dhead_hi) {  // hi/lo pair for the tcgen05 3xTF32 GEMMs of the backward pass
                float* dhh = dhead_hi + row * ld_split;
                float* dhl = dhead_lo + row * ld_split;
#pragma unroll
                for (int j = 0; j < kHead; j++) {
                    const float h = __uint_as_float(__float_as_uint(dv[j]) & 0xFFFFE000u);
                    dhh[j] = h;
                    dhl[j] = dv[j] - h;
                }
            }
            if (vs_out) vs_out[row] = vs;
            if (adv_out) adv_out[row] = adv;
            l_pg += (double)(-logp_a * adv);
            l_bl += 0.5 * (double)(vs - v) * (double)(vs - v);
            l_ent += (double)plogp;
        }
        carry = __shfl_sync(0xffffffffu, acc, 0);
        vs_next_chunk = __shfl_sync(0xffffffffu, vs, 0);
        v_next_chunk = __shfl_sync(0xffffffffu, v, 0);
    }
    l_pg = warp_sum(l_pg);
    l_bl = warp_sum(l_bl);
    l_ent = warp_sum(l_ent);
    if (lane == 0) {
        atomicAdd(losses + 1, l_pg);
        atomicAdd(losses + 2, l_bl);
        atomicAdd(losses + 3, l_ent);
        atomicAdd(losses + 0, l_pg + (double)baseline_cost * l_bl + (double)entropy_cost * l_ent);
    }
}

int launch_vtrace_loss_head(const void* batch, int m, int t, const float* head, int ldh, float rho_bar,
                            float c_bar, float pg_rho_bar, float lambda_, float baseline_cost,
                            float entropy_cost, float* dhead, float* vs, float* pg_adv, double* losses,
                            cudaStream_t stream, float* dhead_hi, float* dhead_lo, int ld_split) {
    if (m <= 0 || t <= 0) return FI_OK;
    if (!batch || !head || (!dhead && !dhead_hi) || !losses || ldh < kHead || (dhead_hi && (!dhead_lo || ld_split < kHead)))
        return set_error(FI_ERR_ARG, "vtrace loss head: bad argument");
    const int warps = 4;
    // algorithmic traffic per transition: head row in (17 x 4 B) + dhead row out (17 x 4 B) + the record's
    // behaviour logits, action, reward, discount (19 x 4 B) = 212 B (+ 8 B when vs / pg_adv are written)
    LaunchScope ls("vtrace_loss_head_kernel", stream,
                   (212.0 + (vs ? 4.0 : 0.0) + (pg_adv ? 4.0 : 0.0)) * (double)m * t, kWorkBytes);
    vtrace_loss_head_kernel<<<(m + warps - 1) / warps, 32 * warps, 0, stream>>>(
        (const float*)batch, m, t, head, ldh, rho_bar, c_bar, pg_rho_bar, lambda_, baseline_cost, entropy_cost,
        dhead, vs, pg_adv, losses, dhead_hi, dhead_lo, ld_split);
    return ls.done();
}

}  // namespace fi

extern "C" {

int fi_op_vtrace(int m, int t, const float* log_rho, const float* discount, const float* reward, const float* value,
                 const float* bootstrap, float rho_bar, float c_bar, float pg_rho_bar, float lambda_, float* vs,
                 float* pg_adv, void* stream) {
    return fi::launch_vtrace_scan(m, t, log_rho, discount, reward, value, bootstrap, rho_bar, c_bar, pg_rho_bar,
                                  lambda_, vs, pg_adv, (cudaStream_t)stream);
}

int fi_op_vtrace_loss_head(const void* batch, int m, int t, const float* head, int ldh, float rho_bar, float c_bar,
                           float pg_rho_bar, float lambda_, float baseline_cost, float entropy_cost, float* dhead,
                           float* vs, float* pg_adv, double* losses, void* stream) {
    return fi::launch_vtrace_loss_head(batch, m, t, head, ldh, rho_bar, c_bar, pg_rho_bar, lambda_, baseline_cost,
                                       entropy_cost, dhead, vs, pg_adv, losses, (cudaStream_t)stream);
}
}
