// Accuracy check of the branch-free gate functions of csrc/model_farmer.cu (fast_exp / sigmoidf_ / fast_tanh, copied below
// verbatim) against double precision over [-30, 30] plus a small-argument and a large-argument family:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o freeimpala_b200/_build/gate_functions_check tools/gate_functions_check.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ float fast_exp(float x) {
    const float t = x * 1.44269502f;
    const float r = fmaf(x, 1.44269502f, -t) + x * 1.925963033e-8f;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    return e * fmaf(r, 0.693147182f, 1.f);
}
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + fast_exp(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
    const float ax = fabsf(x), x2 = ax * ax;
    const float big = 1.f - __fdividef(2.f, fast_exp(2.f * ax) + 1.f);
    const float poly = fmaf(x2, fmaf(x2, fmaf(x2, 0.0218694885f, -0.0539682540f), 0.133333333f), -0.333333333f);
    const float small = fmaf(ax * x2, poly, ax);
    return copysignf(ax < 0.25f ? small : big, x);
}
__global__ void k(const float* x, int n, float* s, float* t) { int i = blockIdx.x*blockDim.x+threadIdx.x; if (i<n) { s[i]=sigmoidf_(x[i]); t[i]=fast_tanh(x[i]); } }
int main() {
    const int n = 1<<22; float *hx = new float[n], *hs = new float[n], *ht = new float[n];
    for (int i = 0; i < n; i++) { double u = (double)i / n; hx[i] = (float)((u - 0.5) * 60.0); if (i % 7 == 0) hx[i] *= 1e-3f; if (i % 11 == 0) hx[i] *= 10.f; }
    float *dx, *ds, *dt; cudaMalloc(&dx, n*4); cudaMalloc(&ds, n*4); cudaMalloc(&dt, n*4);
    cudaMemcpy(dx, hx, n*4, cudaMemcpyHostToDevice);
    k<<<n/256, 256>>>(dx, n, ds, dt); cudaMemcpy(hs, ds, n*4, cudaMemcpyDeviceToHost); cudaMemcpy(ht, dt, n*4, cudaMemcpyDeviceToHost);
    double ms = 0, mt = 0, mta = 0; int nan = 0;
    for (int i = 0; i < n; i++) {
        double x = hx[i], rs = 1.0 / (1.0 + exp(-x)), rt = tanh(x);
        if (std::isnan(hs[i]) || std::isnan(ht[i])) nan++;
        if (rs > 1e-30) ms = fmax(ms, fabs(hs[i] - rs) / rs);
        if (fabs(rt) > 0) mt = fmax(mt, fabs(ht[i] - rt) / fabs(rt));
        mta = fmax(mta, fabs(ht[i] - rt));
    }
    printf("sigmoid max rel err %.3e  tanh max rel err %.3e  tanh max abs err %.3e  nan %d (%s)\n", ms, mt, mta, nan, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
