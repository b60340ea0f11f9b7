// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// C shim around the UNMODIFIED reference learner step. The reference source is compiled
// from where it lies (/root/reference/cmd/libtorch_bench/main.cpp, passed as
// -DFI_REF_BENCH_MAIN=<path> by oracle/Makefile); nothing is copied into this repo.
// Its main() is renamed so the file can live inside a shared library; everything this
// shim calls is the reference's own code:
//   FarmerLstmModel   main.cpp:14-42     make_optimizer  main.cpp:94-103
//   criterion         main.cpp:105-114   train_step      main.cpp:117-135
//   make_batch        main.cpp:85-91
// Output goes to oracle/_ref/libfi_ref_nn.so (git-ignored, travels to the GPU box).
#define main fi_ref_bench_main
#include FI_REF_BENCH_MAIN
#undef main

#include <cstdint>
#include <cstring>

namespace {
struct RefLearner {
    FarmerLstmModel model;
    std::shared_ptr<torch::optim::Optimizer> opt;
    std::string loss_name;
};
}  // namespace

extern "C" {

int ref_nn_num_threads() { return torch::get_num_threads(); }
void ref_nn_set_num_threads(int n) { torch::set_num_threads(n); }

// Seeded random-init model + optimizer, as main.cpp:199-201 does (the reference does not
// seed; we fix torch::manual_seed so fixtures are reproducible).
void* ref_nn_create(uint64_t seed, const char* opt_name, double lr, const char* loss_name) {
    torch::manual_seed(seed);
    auto* r = new RefLearner();
    r->opt = make_optimizer(opt_name, r->model, lr);
    r->loss_name = loss_name;
    return r;
}
void ref_nn_destroy(void* h) { delete static_cast<RefLearner*>(h); }

int ref_nn_num_tensors(void* h) {
    return static_cast<int>(static_cast<RefLearner*>(h)->model.parameters().size());
}
int64_t ref_nn_tensor_numel(void* h, int i) {
    return static_cast<RefLearner*>(h)->model.parameters()[i].numel();
}
int64_t ref_nn_param_count(void* h) {
    int64_t n = 0;
    for (auto& p : static_cast<RefLearner*>(h)->model.parameters()) n += p.numel();
    return n;
}
// Flat fp32 in model.parameters() order (lstm w_ih,w_hh,b_ih,b_hh, dense1.w,.b ... dense6.w,.b).
void ref_nn_get_params(void* h, float* out) {
    for (auto& p : static_cast<RefLearner*>(h)->model.parameters()) {
        auto c = p.detach().contiguous();
        std::memcpy(out, c.data_ptr<float>(), sizeof(float) * c.numel());
        out += c.numel();
    }
}
void ref_nn_set_params(void* h, const float* in) {
    torch::NoGradGuard g;
    for (auto& p : static_cast<RefLearner*>(h)->model.parameters()) {
        auto src = torch::from_blob(const_cast<float*>(in), p.sizes(), torch::kFloat32);
        p.copy_(src);
        in += p.numel();
    }
}
void ref_nn_get_grads(void* h, float* out) {
    for (auto& p : static_cast<RefLearner*>(h)->model.parameters()) {
        if (p.grad().defined()) {
            auto c = p.grad().contiguous();
            std::memcpy(out, c.data_ptr<float>(), sizeof(float) * c.numel());
        } else {
            std::memset(out, 0, sizeof(float) * p.numel());
        }
        out += p.numel();
    }
}

// Fill z,x,target with the reference's own generator (make_batch, main.cpp:85-91).
void ref_nn_make_batch(uint64_t seed, int B, int T, float* z, float* x, float* target) {
    torch::manual_seed(seed);
    Synthetic s = make_batch(B, T, torch::kCPU);
    std::memcpy(z, s.z.data_ptr<float>(), sizeof(float) * s.z.numel());
    std::memcpy(x, s.x.data_ptr<float>(), sizeof(float) * s.x.numel());
    std::memcpy(target, s.target.data_ptr<float>(), sizeof(float) * s.target.numel());
}

// Loss of the current weights on (z,x,target): criterion(forward(...)) as main.cpp:221.
double ref_nn_loss(void* h, const float* z, const float* x, const float* target, int B, int T) {
    auto* r = static_cast<RefLearner*>(h);
    torch::NoGradGuard g;
    auto tz = torch::from_blob(const_cast<float*>(z), {B, T, 162}, torch::kFloat32);
    auto tx = torch::from_blob(const_cast<float*>(x), {B, 484}, torch::kFloat32);
    auto tt = torch::from_blob(const_cast<float*>(target), {B, 1}, torch::kFloat32);
    return criterion(r->loss_name, r->model.forward(tz, tx), tt).item<double>();
}

// One reference train_step (main.cpp:117-135) on caller-provided inputs. Returns the
// reference's own timing (ms). Gradients stay readable via ref_nn_get_grads afterwards.
double ref_nn_train_step(void* h, const float* z, const float* x, const float* target, int B, int T) {
    auto* r = static_cast<RefLearner*>(h);
    Synthetic s;
    s.z = torch::from_blob(const_cast<float*>(z), {B, T, 162}, torch::kFloat32);
    s.x = torch::from_blob(const_cast<float*>(x), {B, 484}, torch::kFloat32);
    s.target = torch::from_blob(const_cast<float*>(target), {B, 1}, torch::kFloat32);
    return train_step(r->model, s, r->loss_name, *r->opt, torch::kCPU);
}

// Forward only: y[B] (main.cpp:25-37).
void ref_nn_forward(void* h, const float* z, const float* x, int B, int T, float* y) {
    auto* r = static_cast<RefLearner*>(h);
    torch::NoGradGuard g;
    auto tz = torch::from_blob(const_cast<float*>(z), {B, T, 162}, torch::kFloat32);
    auto tx = torch::from_blob(const_cast<float*>(x), {B, 484}, torch::kFloat32);
    auto out = r->model.forward(tz, tx).contiguous();
    std::memcpy(y, out.data_ptr<float>(), sizeof(float) * B);
}

// The reference benchmark loop body (main.cpp:213-217): fresh make_batch + train_step,
// `runs` times after `warmups`; returns mean ms like main.cpp:231.
double ref_nn_bench(void* h, int B, int T, int warmups, int runs) {
    auto* r = static_cast<RefLearner*>(h);
    for (int i = 0; i < warmups; ++i) {
        auto batch = make_batch(B, T, torch::kCPU);
        train_step(r->model, batch, r->loss_name, *r->opt, torch::kCPU);
    }
    double total = 0;
    for (int i = 0; i < runs; ++i) {
        auto batch = make_batch(B, T, torch::kCPU);
        total += train_step(r->model, batch, r->loss_name, *r->opt, torch::kCPU);
    }
    return runs > 0 ? total / runs : 0.0;
}

// The reference CLI itself (main.cpp:138-259), argv passed through.
int ref_nn_cli_main(int argc, char** argv) { return fi_ref_bench_main(argc, argv); }

}  // extern "C"
