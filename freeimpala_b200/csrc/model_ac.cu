// MLP actor-critic V-trace learner step (BASELINE.json north_star; configs[1], configs[3]).
// Trunk = the reference's dense1..5 shapes (cmd/libtorch_bench/main.cpp:17-21) applied per
// transition to the 162-feature observation; one fused head [17, 512]: 16 policy logits + value.
// Parameter order: dense1.w [512,162], dense1.b, dense2..5 .w [512,512], .b, head.w [17,512], head.b.
// The first GEMM reads the observations straight out of the gathered batch (row stride 256
// words = one 1024-byte record): no decode pass, no copy.
#include "learner.cuh"

namespace fi {

int ac_alloc(fi_learner* l, Player* p) {
    const size_t rows = l->cfg.batch_size * l->cfg.entry_size;
    p->act.assign(5, nullptr);
    for (int i = 0; i < 5; i++) FI_CUDA_OK(cudaMalloc((void**)&p->act[i], rows * kHid * sizeof(float)));
    FI_CUDA_OK(cudaMalloc((void**)&p->d_a, rows * kHid * sizeof(float)));
    FI_CUDA_OK(cudaMalloc((void**)&p->d_b, rows * kHid * sizeof(float)));
    FI_CUDA_OK(cudaMalloc((void**)&p->head, rows * kHead * sizeof(float)));
    FI_CUDA_OK(cudaMalloc((void**)&p->dhead, rows * kHead * sizeof(float)));
    size_t ws = 0;
    const int mode = l->cfg.gemm_mode;
    auto upd = [&](size_t b) { if (b > ws) ws = b; };
    upd(gemm_workspace_bytes(mode, 2, kHid, kZDim, (int)rows));
    upd(gemm_workspace_bytes(mode, 2, kHid, kHid, (int)rows));
    upd(gemm_workspace_bytes(mode, 2, kHead, kHid, (int)rows));
    upd(gemm_workspace_bytes(mode, 0, (int)rows, kHid, kHid));
    upd(gemm_workspace_bytes(mode, 1, (int)rows, kHid, kHid));
    p->gemm_ws_bytes = ws;
    if (ws) FI_CUDA_OK(cudaMalloc(&p->gemm_ws, ws));
    p->colsum_ws_bytes = colsum_workspace_bytes((int)rows, kHid);
    FI_CUDA_OK(cudaMalloc(&p->colsum_ws, p->colsum_ws_bytes));
    return FI_OK;
}

void ac_free(Player* p) {
    for (auto a : p->act) if (a) cudaFree(a);
    p->act.clear();
    for (auto a : p->inf_act) if (a) cudaFree(a);
    p->inf_act.clear();
}

static int ac_forward(fi_learner* l, const float* params, const float* in, int ld_in, int rows,
                      float* const* act, float* head, void* ws, size_t ws_bytes, cudaStream_t st) {
    const auto& T = l->tensors;
    const int mode = l->cfg.gemm_mode;
    const float* x = in;
    int ldx = ld_in, k = kZDim;
    for (int layer = 0; layer < 5; layer++) {
        FI_TRY(launch_gemm(mode, 0, rows, kHid, k, x, ldx, params + T[2 * layer].offset, k, act[layer], kHid,
                           params + T[2 * layer + 1].offset, 1, nullptr, 0, ws, ws_bytes, st));
        x = act[layer];
        ldx = kHid;
        k = kHid;
    }
    return launch_gemm(mode, 0, rows, kHead, kHid, x, kHid, params + T[10].offset, kHid, head, kHead,
                       params + T[11].offset, 0, nullptr, 0, ws, ws_bytes, st);
}

int ac_forward_backward(fi_learner* l, Player* p, const float* batch, int m, int t, int /*global_m*/) {
    const auto& T = l->tensors;
    const auto& c = l->cfg;
    const int rows = m * t, mode = c.gemm_mode;
    cudaStream_t st = p->stream;
    FI_TRY(ac_forward(l, p->params, batch, kRecWords, rows, p->act.data(), p->head, p->gemm_ws, p->gemm_ws_bytes, st));
    FI_CUDA_OK(cudaMemsetAsync(p->d_losses, 0, 4 * sizeof(double), st));
    FI_TRY(launch_vtrace_loss_head(batch, m, t, p->head, kHead, c.rho_bar, c.c_bar, c.pg_rho_bar, c.lambda_,
                                   c.baseline_cost, c.entropy_cost, p->dhead, nullptr, nullptr, p->d_losses, st));
    // head: dW = dhead^T act4, db = colsum(dhead), d4 = (dhead Wh) * relu'(act4)
    float* g = p->grads;
    FI_TRY(launch_colsum(p->dhead, kHead, rows, kHead, g + T[11].offset, p->colsum_ws, p->colsum_ws_bytes, st));
    FI_TRY(launch_gemm(mode, 2, kHead, kHid, rows, p->dhead, kHead, p->act[4], kHid, g + T[10].offset, kHid, nullptr,
                       0, nullptr, 0, p->gemm_ws, p->gemm_ws_bytes, st));
    float* d = p->d_a;
    float* d_next = p->d_b;
    FI_TRY(launch_gemm(mode, 1, rows, kHid, kHead, p->dhead, kHead, p->params + T[10].offset, kHid, d, kHid, nullptr,
                       0, p->act[4], kHid, p->gemm_ws, p->gemm_ws_bytes, st));
    for (int layer = 4; layer >= 0; layer--) {
        const float* in = layer == 0 ? batch : p->act[layer - 1];
        const int ld_in = layer == 0 ? kRecWords : kHid, k = layer == 0 ? kZDim : kHid;
        FI_TRY(launch_colsum(d, kHid, rows, kHid, g + T[2 * layer + 1].offset, p->colsum_ws, p->colsum_ws_bytes, st));
        FI_TRY(launch_gemm(mode, 2, kHid, k, rows, d, kHid, in, ld_in, g + T[2 * layer].offset, k, nullptr, 0, nullptr,
                           0, p->gemm_ws, p->gemm_ws_bytes, st));
        if (layer > 0) {
            FI_TRY(launch_gemm(mode, 1, rows, kHid, kHid, d, kHid, p->params + T[2 * layer].offset, kHid, d_next, kHid,
                               nullptr, 0, p->act[layer - 1], kHid, p->gemm_ws, p->gemm_ws_bytes, st));
            float* tmp = d; d = d_next; d_next = tmp;
        }
    }
    return FI_OK;
}

int ac_activation(Player* p, int layer, const float** a, const float** lo) {
    if (layer < 0 || layer >= (int)p->act.size()) return set_error(FI_ERR_ARG, "no such hidden layer %d", layer);
    *a = p->act[layer];
    *lo = nullptr;
    return FI_OK;
}

int ac_infer_alloc(fi_learner* /*l*/, Player* p, size_t rows) {
    for (auto a : p->inf_act) if (a) cudaFree(a);
    p->inf_act.assign(5, nullptr);
    for (int i = 0; i < 5; i++) FI_CUDA_OK(cudaMalloc((void**)&p->inf_act[i], rows * kHid * sizeof(float)));
    return FI_OK;
}

// Batched actor policy inference: the same forward GEMM kernels on the
// published device snapshot of the weights. obs_dev [rows,162] -> out_dev [rows,17].
int ac_infer(fi_learner* l, Player* p, const float* params, const float* obs_dev, size_t rows, float* out_dev,
             cudaStream_t stream) {
    return ac_forward(l, params, obs_dev, kZDim, (int)rows, p->inf_act.data(), out_dev, nullptr, 0, stream);
}

}  // namespace fi
