// Fused optimiser update over the flat fp32 parameter arena.
//
// Replaces torch::optim::Adam/SGD/AdamW::step as configured by the reference's make_optimizer
// (cmd/libtorch_bench/main.cpp:94-103; defaults torch optim/adam.h:20-25: betas (0.9, 0.999),
// eps 1e-8, weight_decay 0 (AdamW: 1e-2), amsgrad off). libtorch walks 16 tensors with ~8
// elementwise ops each; here one kernel reads p, g, m, v and writes p, m, v once:
// 28 B/param of algorithmic traffic (SGD: 12 B/param), HBM-bound.
// Arithmetic follows libtorch's order: m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g*g;
// denom = sqrt(v)/sqrt(1-b2^t) + eps; p -= (lr/(1-b1^t)) * m/denom, with the bias
// corrections evaluated in double on the host exactly as libtorch does.
#include <cmath>
#include <cstring>

#include "fi_internal.cuh"

namespace fi {

struct OptScalars {
    float beta1, beta2, one_minus_beta1, one_minus_beta2;
    float step_size, bc2_sqrt, eps, decay;  // decay = 1 - lr*wd (AdamW), else 1
    float lr, grad_scale;
};

// Optional by-products of the update, so that a learner step needs no copy-engine operation on its stream (copy-engine
// hand-overs jitter by 10-300 us, and with 8 data-parallel ranks every step takes the worst rank's jitter):
//  * snapshot: the updated parameters are also written here (the model store's device snapshot, +4 B/param);
//  * losses_src/dst: block 0 copies the step's four loss sums into mapped pinned host memory (zero-copy read-back).
struct OptExtras {
    float* snapshot;
    const double* losses_src;
    double* losses_dst;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const OptScalars& s) {
    g *= s.grad_scale;
    p *= s.decay;
    m = __fadd_rn(__fmul_rn(s.beta1, m), __fmul_rn(s.one_minus_beta1, g));
    v = __fadd_rn(__fmul_rn(s.beta2, v), __fmul_rn(__fmul_rn(s.one_minus_beta2, g), g));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), s.bc2_sqrt), s.eps);
    p = __fsub_rn(p, __fmul_rn(s.step_size, __fdiv_rn(m, denom)));
}

constexpr int kOptThreads = 256;

template <bool kAdam>
__global__ void __launch_bounds__(kOptThreads)
fused_opt_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                 float* __restrict__ v, size_t n, OptScalars s, OptExtras x) {
    if (x.losses_dst && blockIdx.x == 0 && threadIdx.x < 4) {
        x.losses_dst[threadIdx.x] = x.losses_src[threadIdx.x];
        __threadfence_system();
    }
    float4* s4 = reinterpret_cast<float4*>(x.snapshot);
    const size_t n4 = n / 4;
    const size_t stride = (size_t)gridDim.x * kOptThreads;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (size_t i = (size_t)blockIdx.x * kOptThreads + threadIdx.x; i < n4; i += stride) {
        float4 pp = p4[i];
        const float4 gg = __ldg(g4 + i);
        if constexpr (kAdam) {
            float4 mm = m4[i], vv = v4[i];
            adam_elem(pp.x, gg.x, mm.x, vv.x, s);
            adam_elem(pp.y, gg.y, mm.y, vv.y, s);
            adam_elem(pp.z, gg.z, mm.z, vv.z, s);
            adam_elem(pp.w, gg.w, mm.w, vv.w, s);
            m4[i] = mm;
            v4[i] = vv;
        } else {
            pp.x -= s.lr * (gg.x * s.grad_scale);
            pp.y -= s.lr * (gg.y * s.grad_scale);
            pp.z -= s.lr * (gg.z * s.grad_scale);
            pp.w -= s.lr * (gg.w * s.grad_scale);
        }
        p4[i] = pp;
        if (s4) s4[i] = pp;
    }
    // tail (n % 4 values), handled by the first threads of block 0
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t i = n4 * 4 + threadIdx.x;
        float pp = p[i];
        if constexpr (kAdam) {
            float mm = m[i], vv = v[i];
            adam_elem(pp, g[i], mm, vv, s);
            m[i] = mm;
            v[i] = vv;
        } else {
            pp -= s.lr * (g[i] * s.grad_scale);
        }
        p[i] = pp;
        if (x.snapshot) x.snapshot[i] = pp;
    }
}

// Everything a launch of the optimiser kernel needs, by value: shared by the direct launch and by the CUDA-graph path of
// the learner step, which re-parameterises ONE captured kernel node per step (step count -> bias corrections, snapshot and
// loss read-back slots) with cudaGraphExecKernelNodeSetParams.
int opt_launch_desc(int opt_kind, double lr, int64_t step, size_t n, float* p, const float* g, float* m, float* v, float grad_scale,
                    float* snapshot, const double* losses_src, double* losses_dst, OptLaunchDesc* d) {
    if (!p || !g) return set_error(FI_ERR_ARG, "optimiser: null p/g");
    const bool adam = opt_kind == FI_OPT_ADAM || opt_kind == FI_OPT_ADAMW;
    if (!adam && opt_kind != FI_OPT_SGD) return set_error(FI_ERR_ARG, "optimiser: unknown kind %d", opt_kind);
    if (adam && (!m || !v)) return set_error(FI_ERR_ARG, "optimiser: Adam needs m and v");
    if (adam && step < 1) return set_error(FI_ERR_ARG, "optimiser: step counts from 1");
    if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)snapshot) & 15)
        return set_error(FI_ERR_ARG, "optimiser: arenas must be 16-byte aligned");
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    const double wd = opt_kind == FI_OPT_ADAMW ? 1e-2 : 0.0;
    OptScalars s;
    s.beta1 = (float)b1; s.beta2 = (float)b2;
    s.one_minus_beta1 = (float)(1.0 - b1); s.one_minus_beta2 = (float)(1.0 - b2);
    const double bc1 = 1.0 - std::pow(b1, (double)step), bc2 = 1.0 - std::pow(b2, (double)step);
    s.step_size = adam ? (float)(lr / bc1) : 0.f;
    s.bc2_sqrt = adam ? (float)std::sqrt(bc2) : 1.f;
    s.eps = (float)eps;
    s.decay = (float)(1.0 - lr * wd);
    s.lr = (float)lr;
    s.grad_scale = grad_scale;
    const size_t n4 = n / 4 ? n / 4 : 1;
    size_t blocks = (n4 + kOptThreads - 1) / kOptThreads;
    const size_t cap = (size_t)kNumSMs * 8;  // whole waves of 8 resident CTAs per SM
    if (blocks > cap) blocks = cap;
    static_assert(sizeof(OptScalars) <= sizeof(d->scalars) && sizeof(OptExtras) <= sizeof(d->extras), "OptLaunchDesc storage");
    d->func = adam ? (const void*)fused_opt_kernel<true> : (const void*)fused_opt_kernel<false>;
    d->grid = (unsigned)blocks;
    d->block = kOptThreads;
    d->p = p; d->g = g; d->m = m; d->v = v; d->n = n;
    memcpy(d->scalars, &s, sizeof(s));
    const OptExtras x{snapshot, losses_src, losses_src ? losses_dst : nullptr};
    memcpy(d->extras, &x, sizeof(x));
    d->args[0] = &d->p; d->args[1] = &d->g; d->args[2] = &d->m; d->args[3] = &d->v; d->args[4] = &d->n;
    d->args[5] = d->scalars; d->args[6] = d->extras;
    // algorithmic traffic: read p,g,m,v + write p,m,v = 28 B/param (SGD: 12 B/param)
    d->work_bytes = ((adam ? 28.0 : 12.0) + (snapshot ? 4.0 : 0.0)) * (double)n;
    return FI_OK;
}

int launch_opt(int opt_kind, double lr, int64_t step, size_t n, float* p, const float* g, float* m,
               float* v, float grad_scale, cudaStream_t stream, float* snapshot, const double* losses_src, double* losses_dst) {
    if (n == 0) return FI_OK;
    OptLaunchDesc d;
    FI_TRY(opt_launch_desc(opt_kind, lr, step, n, p, g, m, v, grad_scale, snapshot, losses_src, losses_dst, &d));
    LaunchScope ls("fused_opt_kernel", stream, d.work_bytes, kWorkBytes);
    const cudaError_t e = cudaLaunchKernel(d.func, dim3(d.grid), dim3(d.block), d.args, 0, stream);
    if (e != cudaSuccess) return set_error(FI_ERR_CUDA, "launch of fused_opt_kernel failed: %s", cudaGetErrorString(e));
    return ls.done();
}

}  // namespace fi

extern "C" int fi_op_adam(int opt_kind, double lr, int64_t step, size_t n, float* p, const float* g, float* m,
                          float* v, float grad_scale, void* stream) {
    return fi::launch_opt(opt_kind, lr, step, n, p, g, m, v, grad_scale, (cudaStream_t)stream, nullptr, nullptr, nullptr);
}
