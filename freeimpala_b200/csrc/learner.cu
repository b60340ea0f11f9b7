// Learner host orchestration behind the C ABI: replaces the body of Learner::trainModel
// (reference include/freeimpala/learner.h:32-49), the Learner ctor's ring/model set-up
// (:100-140) and Model / ModelManager (data_structures.h:43-157, 310-481).
//
// One CUDA stream per player (p worker threads step concurrently on distinct players,
// learner.h:160-162), one flat fp32 arena each for parameters, gradients and the two Adam
// moments (reference parameters() order, so the published blob is the raw parameter arena),
// a publication stream that moves each new version to pinned host memory off the critical
// path, and one NCCL communicator per player for the data-parallel gradient all-reduce.
// There is no libtorch and no CPU fallback anywhere in this file.
#include "learner.cuh"

#include <dirent.h>
#include <dlfcn.h>
#include <nccl.h>
#include <sys/stat.h>

#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdlib>

using fi::ModelStore;
using fi::Player;
using fi::PublishTicket;
using fi::set_error;
using fi::TensorSpec;

// ------------------------------------------------------------------------------------------
// NCCL is bound at run time (dlopen): the library has no link-time dependency on it, and a
// process that already loaded a libnccl (e.g. through torch) shares that copy.
namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("FI_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already in the process?
            if (!api.handle) api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
#define FI_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name))
        FI_SYM(GetUniqueId, "ncclGetUniqueId");
        FI_SYM(CommInitRank, "ncclCommInitRank");
        FI_SYM(CommDestroy, "ncclCommDestroy");
        FI_SYM(AllReduce, "ncclAllReduce");
        FI_SYM(GroupStart, "ncclGroupStart");
        FI_SYM(GroupEnd, "ncclGroupEnd");
        FI_SYM(GetErrorString, "ncclGetErrorString");
#undef FI_SYM
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GroupStart &&
                 api.GroupEnd && api.GetErrorString;
    });
    return api;
}
#define FI_NCCL_OK(expr)                                                                          \
    do {                                                                                          \
        ncclResult_t _r = (expr);                                                                 \
        if (_r != ncclSuccess)                                                                    \
            return set_error(FI_ERR_NCCL, "%s failed: %s", #expr, nccl().GetErrorString(_r));     \
    } while (0)

// splitmix64: deterministic weight init, U(+-1/sqrt(fan_in)) like torch::nn's defaults
// (the init need not be reproduced; parity tests load exported weights).
struct SplitMix {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

void build_tensor_table(fi_learner* l) {
    auto add = [&](size_t rows, size_t cols, int fan_in) {
        TensorSpec t;
        t.offset = l->param_count;
        t.rows = rows;
        t.cols = cols;
        t.numel = rows * (cols ? cols : 1);
        t.fan_in = fan_in;
        l->tensors.push_back(t);
        l->param_count += t.numel;
    };
    using namespace fi;
    if (l->cfg.model == FI_MODEL_FARMER_LSTM) {
        // FarmerLstmModel parameters() order (reference cmd/libtorch_bench/main.cpp:16-22):
        // lstm.weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0, dense1..6 weight, bias
        add(4 * kLstmH, kZDim, kLstmH);
        add(4 * kLstmH, kLstmH, kLstmH);
        add(4 * kLstmH, 0, kLstmH);
        add(4 * kLstmH, 0, kLstmH);
        add(kHid, kLstmH + kXDim, kLstmH + kXDim);
        add(kHid, 0, kLstmH + kXDim);
        for (int i = 0; i < 4; i++) {
            add(kHid, kHid, kHid);
            add(kHid, 0, kHid);
        }
        add(1, kHid, kHid);
        add(1, 0, kHid);
    } else {
        add(kHid, kZDim, kZDim);
        add(kHid, 0, kZDim);
        for (int i = 0; i < 4; i++) {
            add(kHid, kHid, kHid);
            add(kHid, 0, kHid);
        }
        add(kHead, kHid, kHid);
        add(kHead, 0, kHid);
    }
    l->arena_elems = (l->param_count + 3) & ~(size_t)3;
}

void CUDART_CB publish_cb(void* arg) {
    PublishTicket* t = static_cast<PublishTicket*>(arg);
    ModelStore* s = t->store;
    {
        std::unique_lock<std::shared_mutex> w(s->rw);  // the reference's shared_ptr swap (:441-451)
        s->published = t->host_index;
        s->buf_version[t->host_index] = t->version;
        s->latest_version.store(t->version);
    }
    {
        std::lock_guard<std::mutex> g(s->cv_mu);
    }
    s->cv.notify_all();
    delete t;
}

// Publication = (1) a device snapshot of the weights taken on the learner stream, (2) its D2H copy and the host-blob
// flip on the publication stream. publish_begin makes the learner stream wait until the next snapshot slot is free and
// returns it; the caller fills dev_snap[slot] on the learner stream (the fused optimiser writes it as a by-product, so the
// step has no copy-engine operation; other callers copy) and calls publish_end.
int publish_begin(Player* p, int* slot) {
    ModelStore& s = p->store;
    std::lock_guard<std::mutex> g(s.mu);
    const int sn = s.next_snap;
    if (s.snap_free_recorded[sn]) FI_CUDA_OK(cudaStreamWaitEvent(p->stream, s.snap_free[sn], 0));
    if (s.infer_recorded[sn]) FI_CUDA_OK(cudaStreamWaitEvent(p->stream, s.infer_done[sn], 0));
    *slot = sn;
    return FI_OK;
}

int publish_end(Player* p, uint64_t version, int sn) {
    ModelStore& s = p->store;
    std::lock_guard<std::mutex> g(s.mu);
    FI_CUDA_OK(cudaEventRecord(s.snap_ready[sn], p->stream));
    s.newest_snap = sn;
    s.next_snap = (sn + 1) % ModelStore::kSnaps;
    FI_CUDA_OK(cudaStreamWaitEvent(s.pub_stream, s.snap_ready[sn], 0));
    const int h = s.next_host;
    s.next_host = (h + 1) % 3;
    FI_CUDA_OK(cudaMemcpyAsync(s.host_buf[h], s.dev_snap[sn], s.bytes, cudaMemcpyDeviceToHost, s.pub_stream));
    FI_CUDA_OK(cudaEventRecord(s.snap_free[sn], s.pub_stream));
    s.snap_free_recorded[sn] = true;
    PublishTicket* t = new PublishTicket{&s, h, version};
    cudaError_t e = cudaLaunchHostFunc(s.pub_stream, publish_cb, t);
    if (e != cudaSuccess) {
        delete t;
        return set_error(FI_ERR_CUDA, "cudaLaunchHostFunc failed: %s", cudaGetErrorString(e));
    }
    return FI_OK;
}

// Enqueue the publication of the weights currently in p->params as `version` (snapshot by a device-to-device copy).
int publish_async(Player* p, uint64_t version) {
    int sn = 0;
    FI_TRY(publish_begin(p, &sn));
    FI_CUDA_OK(cudaMemcpyAsync(p->store.dev_snap[sn], p->params, p->store.bytes, cudaMemcpyDeviceToDevice, p->stream));
    return publish_end(p, version, sn);
}

// Synchronous publication (create / set_params / load): everything idle on return.
int publish_sync(Player* p, uint64_t version) {
    FI_TRY(publish_async(p, version));
    FI_CUDA_OK(cudaStreamSynchronize(p->stream));
    FI_CUDA_OK(cudaStreamSynchronize(p->store.pub_stream));
    return FI_OK;
}

Player* get_player(fi_learner* l, int player) {
    if (!l || player < 0 || player >= (int)l->players.size()) {
        set_error(FI_ERR_ARG, "invalid learner or player index %d", player);
        return nullptr;
    }
    return l->players[player];
}

void free_player(fi_learner* l, Player* p) {
    if (!p) return;
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->store.pub_stream) cudaStreamSynchronize(p->store.pub_stream);
    if (p->infer_stream) cudaStreamSynchronize(p->infer_stream);
    if (p->nccl_comm && nccl().ok) nccl().CommDestroy((ncclComm_t)p->nccl_comm);
    if (l->cfg.model == FI_MODEL_FARMER_LSTM) fi::farmer_free(p);
    else fi::ac_free(p);
    float* dev[] = {p->params, p->grads, p->adam_m, p->adam_v, p->d_a, p->d_b, p->head, p->dhead,
                    p->inf_in, p->inf_x, p->inf_out};
    for (float* d : dev)
        if (d) cudaFree(d);
    for (float* d : p->store.dev_snap)
        if (d) cudaFree(d);
    void* devv[] = {p->d_losses, p->gemm_ws, p->colsum_ws, p->model_ws, p->stage_dev, p->inf_model_ws};
    for (void* d : devv)
        if (d) cudaFree(d);
    void* pinned[] = {p->h_losses, p->stage_host, p->inf_host_in, p->inf_host_x, p->inf_host_out,
                      p->store.host_buf[0], p->store.host_buf[1], p->store.host_buf[2]};
    for (void* h : pinned)
        if (h) cudaFreeHost(h);
    for (cudaEvent_t e : p->loss_ev) if (e) cudaEventDestroy(e);
    if (p->batch_ready) cudaEventDestroy(p->batch_ready);
    for (int i = 0; i < ModelStore::kSnaps; i++) {
        cudaEvent_t evs[] = {p->store.snap_ready[i], p->store.snap_free[i], p->store.infer_done[i]};
        for (cudaEvent_t e : evs)
            if (e) cudaEventDestroy(e);
    }
    if (p->stream) cudaStreamDestroy(p->stream);
    if (p->store.pub_stream) cudaStreamDestroy(p->store.pub_stream);
    if (p->infer_stream) cudaStreamDestroy(p->infer_stream);
    delete p;
}

int create_player(fi_learner* l, int index) {
    Player* p = new Player();
    p->index = index;
    l->players.push_back(p);  // owned by the learner from here on (freed by fi_learner_destroy)
    const size_t abytes = l->arena_elems * sizeof(float);
    FI_CUDA_OK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    FI_CUDA_OK(cudaStreamCreateWithFlags(&p->store.pub_stream, cudaStreamNonBlocking));
    FI_CUDA_OK(cudaStreamCreateWithFlags(&p->infer_stream, cudaStreamNonBlocking));
    float** arenas[] = {&p->params, &p->grads, &p->adam_m, &p->adam_v};
    for (float** a : arenas) {
        FI_CUDA_OK(cudaMalloc((void**)a, abytes));
        FI_CUDA_OK(cudaMemset(*a, 0, abytes));
    }
    FI_CUDA_OK(cudaMalloc((void**)&p->d_losses, 4 * sizeof(double)));
    FI_CUDA_OK(cudaMemset(p->d_losses, 0, 4 * sizeof(double)));
    FI_CUDA_OK(cudaHostAlloc((void**)&p->h_losses, (Player::kLossRing + 1) * 4 * sizeof(double), cudaHostAllocPortable | cudaHostAllocMapped));
    memset(p->h_losses, 0, (Player::kLossRing + 1) * 4 * sizeof(double));
    FI_CUDA_OK(cudaHostGetDevicePointer((void**)&p->h_losses_dev, p->h_losses, 0));
    for (cudaEvent_t& e : p->loss_ev) FI_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    FI_CUDA_OK(cudaEventCreateWithFlags(&p->batch_ready, cudaEventDisableTiming));
    ModelStore& s = p->store;
    s.bytes = l->param_count * sizeof(float);
    for (int i = 0; i < ModelStore::kSnaps; i++) {
        FI_CUDA_OK(cudaMalloc((void**)&s.dev_snap[i], abytes));
        FI_CUDA_OK(cudaEventCreateWithFlags(&s.snap_ready[i], cudaEventDisableTiming));
        FI_CUDA_OK(cudaEventCreateWithFlags(&s.snap_free[i], cudaEventDisableTiming));
        FI_CUDA_OK(cudaEventCreateWithFlags(&s.infer_done[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 3; i++) FI_CUDA_OK(cudaHostAlloc((void**)&s.host_buf[i], abytes, cudaHostAllocPortable));
    // step workspaces
    if (l->cfg.model == FI_MODEL_FARMER_LSTM) FI_TRY(fi::farmer_alloc(l, p));
    else FI_TRY(fi::ac_alloc(l, p));
    // random init, distinct per player (the reference fills each player's Model with rand(), :121-127)
    std::vector<float> init(l->arena_elems, 0.f);
    SplitMix rng{l->cfg.seed * 0x100000001B3ull + (uint64_t)index + 1};
    for (const TensorSpec& t : l->tensors) {
        const double k = 1.0 / std::sqrt((double)t.fan_in);
        for (size_t i = 0; i < t.numel; i++) init[t.offset + i] = (float)((2.0 * rng.uniform() - 1.0) * k);
    }
    FI_CUDA_OK(cudaMemcpy(p->params, init.data(), abytes, cudaMemcpyHostToDevice));
    p->version = 1;  // Model ctor: version 0, generateRandomData() -> 1 (data_structures.h:52-58,121-127)
    s.next_snap = 0;
    s.next_host = 0;
    return publish_sync(p, p->version);
}

// mkdir -p
bool make_dirs(const std::string& path) {
    if (path.empty()) return false;
    for (size_t i = 1; i <= path.size(); i++) {
        if (i != path.size() && path[i] != '/') continue;
        const std::string part = path.substr(0, i);
        if (mkdir(part.c_str(), 0777) != 0 && errno != EEXIST) return false;
    }
    struct stat st;
    return stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

bool file_exists(const std::string& path) {
    struct stat st;
    return stat(path.c_str(), &st) == 0;
}

const char kOptMagic[8] = {'F', 'I', 'O', 'P', 'T', '0', '0', '1'};

}  // namespace

extern "C" {

const char* fi_last_error(void) { return fi::last_error_ref().c_str(); }
const char* fi_version(void) { return "freeimpala-b200 0.1 (sm_100a)"; }
uint64_t fi_kernel_launch_count(void) { return fi::launch_counter().load(); }

void fi_learner_config_default(fi_learner_config* c) {
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->device = 0;
    c->num_players = 2;          // -p (reference cmd/freeimpala/main.cpp:38-120 defaults)
    c->buffer_capacity = 10;     // -B
    c->entry_size = 100;         // -S
    c->batch_size = 5;           // -M
    c->model = FI_MODEL_MLP_ACTOR_CRITIC;
    c->loss = FI_LOSS_VTRACE;
    c->optimizer = FI_OPT_ADAM;
    c->lr = 5e-4;                // README bench shape
    c->seed = 0;
    c->rho_bar = 1.f; c->c_bar = 1.f; c->pg_rho_bar = 1.f; c->lambda_ = 1.f;
    c->baseline_cost = 0.5f; c->entropy_cost = 0.01f;
    c->gemm_mode = FI_GEMM_AUTO;
    c->publish_every = 1;
    c->checkpoint_location = nullptr;
}

fi_learner* fi_learner_create(const fi_learner_config* cfg) {
    if (!cfg || cfg->num_players < 1 || cfg->entry_size == 0 || cfg->batch_size == 0 ||
        cfg->buffer_capacity < cfg->batch_size) {  // validateParameters: M <= B (main.cpp:160-172)
        set_error(FI_ERR_ARG, "fi_learner_create: invalid configuration (need p>=1, S>0, 0<M<=B)");
        return nullptr;
    }
    const bool farmer = cfg->model == FI_MODEL_FARMER_LSTM;
    if ((farmer && cfg->loss == FI_LOSS_VTRACE) || (!farmer && cfg->loss != FI_LOSS_VTRACE) ||
        (cfg->model != FI_MODEL_FARMER_LSTM && cfg->model != FI_MODEL_MLP_ACTOR_CRITIC)) {
        set_error(FI_ERR_ARG, "fi_learner_create: model %d does not support loss %d", cfg->model, cfg->loss);
        return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error(FI_ERR_CUDA, "fi_learner_create: no CUDA device (there is no CPU fallback)");
        return nullptr;
    }
    FI_CUDA_OK_NULL(cudaSetDevice(cfg->device));
    fi_learner* l = new fi_learner();
    l->cfg = *cfg;
    if (l->cfg.publish_every < 1) l->cfg.publish_every = 1;
    if (cfg->checkpoint_location) l->ckpt_dir = cfg->checkpoint_location;
    l->cfg.checkpoint_location = nullptr;
    build_tensor_table(l);
    for (int p = 0; p < cfg->num_players; p++) {
        fi_ring* r = fi_ring_create(cfg->device, cfg->entry_size, cfg->buffer_capacity);  // learner.h:135-139
        if (!r) {
            fi_learner_destroy(l);
            return nullptr;
        }
        l->rings.push_back(r);
        if (create_player(l, p) < 0) {
            fi_learner_destroy(l);
            return nullptr;
        }
    }
    return l;
}

void fi_learner_destroy(fi_learner* l) {
    if (!l) return;
    cudaSetDevice(l->cfg.device);
    for (Player* p : l->players) free_player(l, p);
    for (fi_ring* r : l->rings) fi_ring_destroy(r);
    delete l;
}

fi_ring* fi_learner_ring(fi_learner* l, int player) {
    if (!l || player < 0 || player >= (int)l->rings.size()) return nullptr;
    return l->rings[player];
}
void* fi_learner_stream(fi_learner* l, int player) {
    Player* p = get_player(l, player);
    return p ? (void*)p->stream : nullptr;
}

int fi_learner_sync(fi_learner* l, int player) {
    Player* p = get_player(l, player);
    if (!p) return FI_ERR_ARG;
    FI_CUDA_OK(cudaSetDevice(l->cfg.device));
    FI_CUDA_OK(cudaStreamSynchronize(p->stream));
    FI_CUDA_OK(cudaStreamSynchronize(p->store.pub_stream));
    return FI_OK;
}

int fi_learner_stage_batch(fi_learner* l, int player, const void* host, size_t num_slots, fi_batch* out) {
    Player* p = get_player(l, player);
    if (!p || !host || !out) return set_error(FI_ERR_ARG, "fi_learner_stage_batch: null argument");
    if (num_slots == 0 || num_slots > l->cfg.batch_size)
        return set_error(FI_ERR_ARG, "fi_learner_stage_batch: num_slots %zu not in [1, batch_size=%zu]", num_slots,
                         l->cfg.batch_size);
    FI_CUDA_OK(cudaSetDevice(l->cfg.device));
    const size_t slot_bytes = l->cfg.entry_size * FI_ELEMENT_SIZE, cap = l->cfg.batch_size * slot_bytes;
    const size_t n = num_slots * slot_bytes;
    if (!p->stage_dev) {
        FI_CUDA_OK(cudaMalloc((void**)&p->stage_dev, cap));
        p->stage_bytes = cap;
    }
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const void* src = host;
    if (!pinned) {  // pageable source: bounce through a pinned buffer so the copy is a real async DMA
        if (!p->stage_host) FI_CUDA_OK(cudaHostAlloc((void**)&p->stage_host, cap, cudaHostAllocPortable));
        FI_CUDA_OK(cudaStreamSynchronize(p->stream));  // previous DMA out of the bounce buffer
        memcpy(p->stage_host, host, n);
        src = p->stage_host;
    }
    FI_CUDA_OK(cudaMemcpyAsync(p->stage_dev, src, n, cudaMemcpyHostToDevice, p->stream));
    out->dev_ptr = p->stage_dev;
    out->num_slots = num_slots;
    out->slot_bytes = slot_bytes;
    out->stream = p->stream;
    out->seq = 0;
    return FI_OK;
}

int fi_learner_forward_backward(fi_learner* l, int player, const fi_batch* b) {
    Player* p = get_player(l, player);
    if (!p || !b) return set_error(FI_ERR_ARG, "fi_learner_forward_backward: null argument");
    if (b->num_slots == 0 || b->num_slots > l->cfg.batch_size || b->slot_bytes != l->cfg.entry_size * FI_ELEMENT_SIZE ||
        !b->dev_ptr)
        return set_error(FI_ERR_ARG, "fi_learner_forward_backward: batch [%zu x %zu B] does not match the learner (M<=%zu, S=%zu)",
                         b->num_slots, b->slot_bytes, l->cfg.batch_size, l->cfg.entry_size);
    FI_CUDA_OK(cudaSetDevice(l->cfg.device));
    std::lock_guard<std::mutex> step_lock(p->step_mu);
    if (b->stream && (cudaStream_t)b->stream != p->stream) {  // gather ran on another stream
        FI_CUDA_OK(cudaEventRecord(p->batch_ready, (cudaStream_t)b->stream));
        FI_CUDA_OK(cudaStreamWaitEvent(p->stream, p->batch_ready, 0));
    }
    const int m = (int)b->num_slots, t = (int)l->cfg.entry_size;
    // mean losses divide by the GLOBAL batch: under data parallelism that is the sum of the ranks' configured batch sizes
    // (all-reduced once in fi_learner_dp_init; shards may differ by one trajectory), so every rank must step full batches
    if (l->dp_world > 1 && (size_t)m != l->cfg.batch_size)
        return set_error(FI_ERR_STATE, "fi_learner_forward_backward: a partial batch (%d of %zu trajectories) cannot be combined with "
                                       "data parallelism (the global batch size is fixed at fi_learner_dp_init)", m, l->cfg.batch_size);
    const int global_m = l->dp_world > 1 ? (int)l->dp_global_batch : m;
    if (l->cfg.model == FI_MODEL_FARMER_LSTM) FI_TRY(fi::farmer_forward_backward(l, p, (const float*)b->dev_ptr, m, t, global_m));
    else FI_TRY(fi::ac_forward_backward(l, p, (const float*)b->dev_ptr, m, t, global_m));
    p->last_rows = (size_t)m * t;
    p->grads_valid = true;
    // the batch buffer may be rewritten by its ring's next gather once everything enqueued above has read it
    FI_TRY(fi::ring_note_consumed(b->dev_ptr, p->stream));
    return FI_OK;
}

int fi_learner_apply_update(fi_learner* l, int player) {
    Player* p = get_player(l, player);
    if (!p) return FI_ERR_ARG;
    FI_CUDA_OK(cudaSetDevice(l->cfg.device));
    std::lock_guard<std::mutex> step_lock(p->step_mu);
    if (l->dp_world > 1) {  // sum-allreduce of the flat gradient arena over NVLink
        if (!p->nccl_comm) return set_error(FI_ERR_STATE, "data parallelism configured but the communicator is missing");
        NcclApi& n = nccl();
        // timed as "nccl_allreduce_grads" (bus bytes of a ring all-reduce: 2(N-1)/N x the arena); it also absorbs the wait
        // for the slowest rank, so it reads as skew + transfer
        fi::LaunchScope ls("nccl_allreduce_grads", p->stream,
                           2.0 * (l->dp_world - 1) / l->dp_world * 4.0 * (double)l->param_count, fi::kWorkBytes);
        FI_NCCL_OK(n.GroupStart());
        FI_NCCL_OK(n.AllReduce(p->grads, p->grads, l->param_count, ncclFloat, ncclSum, (ncclComm_t)p->nccl_comm, p->stream));
        FI_NCCL_OK(n.AllReduce(p->d_losses, p->d_losses, 4, ncclDouble, ncclSum, (ncclComm_t)p->nccl_comm, p->stream));
        FI_NCCL_OK(n.GroupEnd());
        ls.done_external();
    }
    p->opt_step++;
    const uint64_t steps_done = p->steps_done.load(std::memory_order_relaxed) + 1;
    p->version++;  // generateRandomData(): version++ (data_structures.h:121-127), then updateModel
    const bool publish = steps_done % (uint64_t)l->cfg.publish_every == 0;
    int sn = -1;
    if (publish) FI_TRY(publish_begin(p, &sn));
    // one kernel: the update, the model store's device snapshot and the loss read-back (into mapped pinned memory)
    const int slot = (int)(steps_done % Player::kLossRing);
    FI_TRY(fi::launch_opt(l->cfg.optimizer, l->cfg.lr, p->opt_step, l->param_count, p->params, p->grads, p->adam_m,
                          p->adam_v, 1.0f, p->stream, publish ? p->store.dev_snap[sn] : nullptr, p->d_losses,
                          p->h_losses_dev + 4 * slot));
    FI_CUDA_OK(cudaEventRecord(p->loss_ev[slot], p->stream));
    p->steps_done.store(steps_done, std::memory_order_release);   // readers (losses_at, steps_done) see the event recorded
    if (publish) FI_TRY(publish_end(p, p->version, sn));
    return FI_OK;
}

int fi_learner_step(fi_learner* l, int player, const fi_batch* batch) {
    FI_TRY(fi_learner_forward_backward(l, player, batch));
    return fi_learner_apply_update(l, player);
}

int fi_learner_last_losses(fi_learner* l, int player, float losses[4]) {
    Player* p = get_player(l, player);
    if (!p || !losses) return set_error(FI_ERR_ARG, "fi_learner_last_losses: null argument");
    FI_CUDA_OK(cudaSetDevice(l->cfg.device));
    const uint64_t done = p->steps_done.load(std::memory_order_acquire);
    const int slot = (int)(done % Player::kLossRing);
    if (done == 0 && p->grads_valid) {  // forward_backward only: read straight from the device
        FI_CUDA_OK(cudaMemcpyAsync(p->h_losses, p->d_losses, 4 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
        FI_CUDA_OK(cudaEventRecord(p->loss_ev[0], p->stream));
    }
    FI_CUDA_OK(cudaEventSynchronize(p->loss_ev[slot]));
    for (int i = 0; i < 4; i++) losses[i] = (float)p->h_losses[4 * slot + i];
    return FI_OK;
}
int fi_learner_losses_at(fi_learner* l, int player, uint64_t step, float losses[4]) {
    Player* p = get_player(l, player);
    if (!p || !losses) return set_error(FI_ERR_ARG, "fi_learner_losses_at: null argument");
    const uint64_t done = p->steps_done.load(std::memory_order_acquire);   // the worker thread increments it under step_mu
    if (step == 0 || step > done || step + Player::kLossRing <= done)
        return set_error(FI_ERR_ARG, "fi_learner_losses_at: step %llu is not among the last %d of %llu", (unsigned long long)step,
                         Player::kLossRing, (unsigned long long)done);
    FI_CUDA_OK(cudaSetDevice(l->cfg.device));
    const int slot = (int)(step % Player::kLossRing);
    FI_CUDA_OK(cudaEventSynchronize(p->loss_ev[slot]));
    for (int i = 0; i < 4; i++) losses[i] = (float)p->h_losses[4 * slot + i];
    return FI_OK;
}
int fi_learner_last_losses_f64(fi_learner* l, int player, double losses[4]) {
    Player* p = get_player(l, player);
    if (!p || !losses) return set_error(FI_ERR_ARG, "fi_learner_last_losses_f64: null argument");
    FI_CUDA_OK(cudaSetDevice(l->cfg.device));
    FI_CUDA_OK(cudaMemcpyAsync(p->h_losses + 4 * Player::kLossRing, p->d_losses, 4 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    FI_CUDA_OK(cudaStreamSynchronize(p->stream));
    for (int i = 0; i < 4; i++) losses[i] = p->h_losses[4 * Player::kLossRing + i];  // scratch slot behind the ring
    return FI_OK;
}
uint64_t fi_learner_steps_done(fi_learner* l, int player) {
    Player* p = get_player(l, player);
    return p ? p->steps_done.load(std::memory_order_acquire) : 0;
}

size_t fi_learner_param_count(const fi_learner* l) { return l ? l->param_count : 0; }
int fi_learner_num_tensors(const fi_learner* l) { return l ? (int)l->tensors.size() : 0; }
int fi_learner_tensor_info(const fi_learner* l, int i, size_t* offset, size_t* numel, size_t* rows, size_t* cols) {
    if (!l || i < 0 || i >= (int)l->tensors.size()) return set_error(FI_ERR_ARG, "fi_learner_tensor_info: bad index");
    const TensorSpec& t = l->tensors[i];
    if (offset) *offset = t.offset;
    if (numel) *numel = t.numel;
    if (rows) *rows = t.rows;
    if (cols) *cols = t.cols;
    return FI_OK;
}

static int arena_io(fi_learner* l, int player, float* host_out, const float* host_in, size_t n, int which) {
    Player* p = get_player(l, player);
    if (!p || (!host_out && !host_in)) return set_error(FI_ERR_ARG, "arena access: null argument");
    if (n != l->param_count) return set_error(FI_ERR_ARG, "arena access: n=%zu but the model has %zu parameters", n, l->param_count);
    FI_CUDA_OK(cudaSetDevice(l->cfg.device));
    std::lock_guard<std::mutex> step_lock(p->step_mu);
    float* arena = which == 0 ? p->params : which == 1 ? p->grads : which == 2 ? p->adam_m : p->adam_v;
    FI_CUDA_OK(cudaStreamSynchronize(p->stream));
    if (host_out) FI_CUDA_OK(cudaMemcpy(host_out, arena, n * sizeof(float), cudaMemcpyDeviceToHost));
    else FI_CUDA_OK(cudaMemcpy(arena, host_in, n * sizeof(float), cudaMemcpyHostToDevice));
    if (host_in && which == 0) FI_TRY(publish_sync(p, p->version));  // actors see the loaded weights
    return FI_OK;
}
int fi_learner_set_params(fi_learner* l, int player, const float* host, size_t n) { return arena_io(l, player, nullptr, host, n, 0); }
int fi_learner_get_params(fi_learner* l, int player, float* host, size_t n) { return arena_io(l, player, host, nullptr, n, 0); }
int fi_learner_get_grads(fi_learner* l, int player, float* host, size_t n) { return arena_io(l, player, host, nullptr, n, 1); }
int fi_learner_set_grads(fi_learner* l, int player, const float* host, size_t n) { return arena_io(l, player, nullptr, host, n, 1); }
int fi_learner_get_opt_state(fi_learner* l, int player, float* m, float* v, size_t n, int64_t* step) {
    Player* p = get_player(l, player);
    if (!p) return FI_ERR_ARG;
    if (m) FI_TRY(arena_io(l, player, m, nullptr, n, 2));
    if (v) FI_TRY(arena_io(l, player, v, nullptr, n, 3));
    if (step) *step = p->opt_step;
    return FI_OK;
}
void* fi_learner_grad_ptr(fi_learner* l, int player) {
    Player* p = get_player(l, player);
    return p ? p->grads : nullptr;
}
void* fi_learner_param_ptr(fi_learner* l, int player) {
    Player* p = get_player(l, player);
    return p ? p->params : nullptr;
}

// ---- batched actor policy inference --------------------------------
int fi_learner_infer(fi_learner* l, int player, const float* obs_or_z, const float* x, size_t rows, size_t t,
                     float* logits, float* values) {
    Player* p = get_player(l, player);
    if (!p || !obs_or_z || rows == 0) return set_error(FI_ERR_ARG, "fi_learner_infer: null argument");
    const bool farmer = l->cfg.model == FI_MODEL_FARMER_LSTM;
    if (farmer && (!x || t == 0 || !values)) return set_error(FI_ERR_ARG, "fi_learner_infer: the farmer model needs z, x, t and values");
    FI_CUDA_OK(cudaSetDevice(l->cfg.device));
    std::lock_guard<std::mutex> infer_lock(p->infer_mu);
    const size_t in_elems = farmer ? rows * t * fi::kZDim : rows * fi::kZDim;
    const size_t out_cols = farmer ? 1 : fi::kHead;
    if (rows > p->inf_rows_cap || (farmer && t > p->inf_t_cap)) {  // grow the inference workspaces
        FI_CUDA_OK(cudaStreamSynchronize(p->infer_stream));
        if (p->inf_in) cudaFree(p->inf_in);
        if (p->inf_x) cudaFree(p->inf_x);
        if (p->inf_out) cudaFree(p->inf_out);
        if (p->inf_host_in) cudaFreeHost(p->inf_host_in);
        if (p->inf_host_x) cudaFreeHost(p->inf_host_x);
        if (p->inf_host_out) cudaFreeHost(p->inf_host_out);
        p->inf_in = p->inf_x = p->inf_out = p->inf_host_in = p->inf_host_x = p->inf_host_out = nullptr;
        p->inf_rows_cap = 0;
        FI_CUDA_OK(cudaMalloc((void**)&p->inf_in, in_elems * sizeof(float)));
        FI_CUDA_OK(cudaHostAlloc((void**)&p->inf_host_in, in_elems * sizeof(float), cudaHostAllocPortable));
        FI_CUDA_OK(cudaMalloc((void**)&p->inf_out, rows * out_cols * sizeof(float)));
        FI_CUDA_OK(cudaHostAlloc((void**)&p->inf_host_out, rows * out_cols * sizeof(float), cudaHostAllocPortable));
        if (farmer) {
            FI_CUDA_OK(cudaMalloc((void**)&p->inf_x, rows * fi::kXDim * sizeof(float)));
            FI_CUDA_OK(cudaHostAlloc((void**)&p->inf_host_x, rows * fi::kXDim * sizeof(float), cudaHostAllocPortable));
            FI_TRY(fi::farmer_infer_alloc(l, p, rows, t));
        } else {
            FI_TRY(fi::ac_infer_alloc(l, p, rows));
        }
        p->inf_rows_cap = rows;
        p->inf_t_cap = t;
    }
    cudaStream_t st = p->infer_stream;
    memcpy(p->inf_host_in, obs_or_z, in_elems * sizeof(float));
    FI_CUDA_OK(cudaMemcpyAsync(p->inf_in, p->inf_host_in, in_elems * sizeof(float), cudaMemcpyHostToDevice, st));
    if (farmer) {
        memcpy(p->inf_host_x, x, rows * fi::kXDim * sizeof(float));
        FI_CUDA_OK(cudaMemcpyAsync(p->inf_x, p->inf_host_x, rows * fi::kXDim * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    {
        ModelStore& s = p->store;
        std::lock_guard<std::mutex> g(s.mu);  // pins the newest snapshot against the learner's next D2D
        const int sn = s.newest_snap;
        FI_CUDA_OK(cudaStreamWaitEvent(st, s.snap_ready[sn], 0));
        if (farmer) FI_TRY(fi::farmer_infer(l, p, s.dev_snap[sn], p->inf_in, p->inf_x, rows, t, p->inf_out, st));
        else FI_TRY(fi::ac_infer(l, p, s.dev_snap[sn], p->inf_in, rows, p->inf_out, st));
        FI_CUDA_OK(cudaEventRecord(s.infer_done[sn], st));
        s.infer_recorded[sn] = true;
    }
    FI_CUDA_OK(cudaMemcpyAsync(p->inf_host_out, p->inf_out, rows * out_cols * sizeof(float), cudaMemcpyDeviceToHost, st));
    FI_CUDA_OK(cudaStreamSynchronize(st));
    if (farmer) {
        memcpy(values, p->inf_host_out, rows * sizeof(float));
    } else {
        for (size_t r = 0; r < rows; r++) {
            const float* o = p->inf_host_out + r * fi::kHead;
            if (logits) memcpy(logits + r * fi::kNumActions, o, fi::kNumActions * sizeof(float));
            if (values) values[r] = o[fi::kNumActions];
        }
    }
    return FI_OK;
}

// ---- model store -----------------------------------------------------------------------------
size_t fi_model_bytes(const fi_learner* l) { return l ? l->param_count * sizeof(float) : 0; }

uint64_t fi_model_version(fi_learner* l, int player) {
    Player* p = get_player(l, player);
    return p ? p->store.latest_version.load() : 0;  // getLatestVersion: 0 for a bad index (:475-480)
}

int fi_model_get(fi_learner* l, int player, void* dst, size_t n, uint64_t* version) {
    Player* p = get_player(l, player);
    if (!p || !dst) return set_error(FI_ERR_ARG, "fi_model_get: null argument");
    ModelStore& s = p->store;
    if (n < s.bytes) return set_error(FI_ERR_ARG, "fi_model_get: buffer of %zu B is smaller than the %zu B blob", n, s.bytes);
    std::shared_lock<std::shared_mutex> r(s.rw);
    memcpy(dst, s.host_buf[s.published], s.bytes);
    if (version) *version = s.buf_version[s.published];
    return FI_OK;
}

int fi_model_wait_update(fi_learner* l, int player, uint64_t current_version, int timeout_ms) {
    Player* p = get_player(l, player);
    if (!p) return 0;
    ModelStore& s = p->store;
    std::unique_lock<std::mutex> lock(s.cv_mu);
    if (s.latest_version.load() > current_version) return 1;
    return s.cv.wait_for(lock, std::chrono::milliseconds(timeout_ms),
                         [&] { return s.latest_version.load() > current_version; }) ? 1 : 0;
}

int fi_model_save(fi_learner* l, int player, uint64_t iteration, int with_optimizer_state) {
    Player* p = get_player(l, player);
    if (!p) return FI_ERR_ARG;
    if (l->ckpt_dir.empty()) return set_error(FI_ERR_IO, "fi_model_save: no checkpoint location configured");
    FI_CUDA_OK(cudaSetDevice(l->cfg.device));
    ModelStore& s = p->store;
    std::vector<unsigned char> blob(s.bytes);
    std::vector<float> om, ov;
    uint64_t version = 0, opt_step = 0;
    if (with_optimizer_state) {  // consistent (weights, m, v, step): quiesce the player's stream
        std::lock_guard<std::mutex> step_lock(p->step_mu);
        FI_CUDA_OK(cudaStreamSynchronize(p->stream));
        om.resize(l->param_count);
        ov.resize(l->param_count);
        FI_CUDA_OK(cudaMemcpy(blob.data(), p->params, s.bytes, cudaMemcpyDeviceToHost));
        FI_CUDA_OK(cudaMemcpy(om.data(), p->adam_m, s.bytes, cudaMemcpyDeviceToHost));
        FI_CUDA_OK(cudaMemcpy(ov.data(), p->adam_v, s.bytes, cudaMemcpyDeviceToHost));
        version = p->version;
        opt_step = (uint64_t)p->opt_step;
    } else {  // the published model, as the reference's saveModel copies models[p] (:396)
        std::shared_lock<std::shared_mutex> r(s.rw);
        memcpy(blob.data(), s.host_buf[s.published], s.bytes);
        version = s.buf_version[s.published];
    }
    uint64_t stamp = iteration;
    if (iteration == 0) {  // :405
        std::lock_guard<std::mutex> g(s.mu);
        stamp = p->checkpoint_counter++;
    }
    // create_directories (data_structures.h:94-97): every missing component of the path
    if (!make_dirs(l->ckpt_dir)) return set_error(FI_ERR_IO, "fi_model_save: cannot create directory %s", l->ckpt_dir.c_str());
    const std::string base = l->ckpt_dir + "/model_" + std::to_string(player) + "_";
    const std::string paths[2] = {base + std::to_string(stamp) + ".bin", base + "latest.bin"};
    // One save per player at a time, and every file appears under its final name only when it is complete (written to a
    // temporary name in the same directory, then rename()d): a checkpoint thread and the final save of stop() can no
    // longer interleave their writes into one latest.bin.
    std::lock_guard<std::mutex> save_lock(p->save_mu);
    for (const std::string& final_path : paths) {
        const std::string path = final_path + ".tmp";
        FILE* f = fopen(path.c_str(), "wb");
        if (!f) return set_error(FI_ERR_IO, "fi_model_save: cannot open %s", path.c_str());
        bool ok = fwrite(&version, sizeof(version), 1, f) == 1 && fwrite(blob.data(), 1, blob.size(), f) == blob.size();
        if (ok && with_optimizer_state) {
            const uint64_t n = l->param_count;
            ok = fwrite(kOptMagic, 1, 8, f) == 8 && fwrite(&opt_step, 8, 1, f) == 1 && fwrite(&n, 8, 1, f) == 1 &&
                 fwrite(om.data(), 4, n, f) == n && fwrite(ov.data(), 4, n, f) == n;
        }
        ok = (fclose(f) == 0) && ok;
        if (!ok) {
            remove(path.c_str());
            return set_error(FI_ERR_IO, "fi_model_save: short write to %s", path.c_str());
        }
        if (rename(path.c_str(), final_path.c_str()) != 0)
            return set_error(FI_ERR_IO, "fi_model_save: cannot rename %s to %s", path.c_str(), final_path.c_str());
    }
    return FI_OK;
}

int fi_model_load(fi_learner* l, const char* dir) {
    if (!l) return FI_ERR_ARG;
    if (!dir || !*dir) return 0;  // loadModels: empty path is a no-op (:338)
    FI_CUDA_OK(cudaSetDevice(l->cfg.device));
    int loaded = 0;
    for (int pi = 0; pi < (int)l->players.size(); pi++) {
        Player* p = l->players[pi];
        const std::string prefix = "model_" + std::to_string(pi) + "_";
        std::string path = std::string(dir) + "/" + prefix + "latest.bin";
        if (!file_exists(path)) {  // highest-numbered model_{p}_N.bin (:344-376)
            uint64_t best = 0;
            std::string best_file;
            if (DIR* d = opendir(dir)) {
                while (dirent* e = readdir(d)) {
                    const std::string name = e->d_name;
                    if (name.compare(0, prefix.size(), prefix) != 0) continue;
                    const size_t end = name.find(".bin");
                    if (end == std::string::npos) continue;
                    const std::string num = name.substr(prefix.size(), end - prefix.size());
                    if (num.empty() || num.find_first_not_of("0123456789") != std::string::npos) continue;
                    const uint64_t it = strtoull(num.c_str(), nullptr, 10);
                    if (it > best) {
                        best = it;
                        best_file = std::string(dir) + "/" + name;
                    }
                }
                closedir(d);
            }
            if (best_file.empty()) continue;
            path = best_file;
            p->checkpoint_counter = best + 1;  // :373
        }
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) continue;
        uint64_t version = 0;
        std::vector<float> blob(l->param_count);
        bool ok = fread(&version, 8, 1, f) == 1 && fread(blob.data(), 4, l->param_count, f) == l->param_count;
        if (!ok) {  // loadFromDisk returns false on a short file and the model keeps its state (:75-84)
            fclose(f);
            continue;
        }
        char magic[8];
        uint64_t opt_step = 0, n = 0;
        std::vector<float> om, ov;
        bool have_opt = fread(magic, 1, 8, f) == 8 && memcmp(magic, kOptMagic, 8) == 0 && fread(&opt_step, 8, 1, f) == 1 &&
                        fread(&n, 8, 1, f) == 1 && n == l->param_count;
        if (have_opt) {
            om.resize(n);
            ov.resize(n);
            have_opt = fread(om.data(), 4, n, f) == n && fread(ov.data(), 4, n, f) == n;
        }
        fclose(f);
        std::lock_guard<std::mutex> step_lock(p->step_mu);
        FI_CUDA_OK(cudaStreamSynchronize(p->stream));
        FI_CUDA_OK(cudaMemcpy(p->params, blob.data(), l->param_count * 4, cudaMemcpyHostToDevice));
        if (have_opt) {
            FI_CUDA_OK(cudaMemcpy(p->adam_m, om.data(), n * 4, cudaMemcpyHostToDevice));
            FI_CUDA_OK(cudaMemcpy(p->adam_v, ov.data(), n * 4, cudaMemcpyHostToDevice));
            p->opt_step = (int64_t)opt_step;
        }
        p->version = version;
        FI_TRY(publish_sync(p, version));
        loaded++;
    }
    return loaded;
}

// ---- data parallelism ------------------------------------------------------------------------
int fi_dp_create_id(void* id_out) {
    if (!id_out) return set_error(FI_ERR_ARG, "fi_dp_create_id: null argument");
    if (!nccl().ok) return set_error(FI_ERR_NCCL, "NCCL library not found (set FI_NCCL_LIB)");
    static_assert(sizeof(ncclUniqueId) == FI_DP_ID_BYTES, "id size");
    ncclUniqueId id;
    FI_NCCL_OK(nccl().GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return FI_OK;
}

int fi_learner_dp_init(fi_learner* l, const void* ids, int rank, int world_size) {
    if (!l || !ids || world_size < 1 || rank < 0 || rank >= world_size)
        return set_error(FI_ERR_ARG, "fi_learner_dp_init: bad arguments");
    if (l->dp_world > 1) return set_error(FI_ERR_STATE, "fi_learner_dp_init: already initialised");
    if (!nccl().ok) return set_error(FI_ERR_NCCL, "NCCL library not found (set FI_NCCL_LIB)");
    FI_CUDA_OK(cudaSetDevice(l->cfg.device));
    for (size_t pi = 0; pi < l->players.size(); pi++) {
        ncclUniqueId id;
        memcpy(&id, (const char*)ids + pi * FI_DP_ID_BYTES, sizeof(id));
        ncclComm_t comm = nullptr;
        FI_NCCL_OK(nccl().CommInitRank(&comm, world_size, id, rank));
        l->players[pi]->nccl_comm = comm;
    }
    l->dp_rank = rank;
    l->dp_world = world_size;
    // the global batch = sum of the ranks' batch sizes (shards of a global batch may differ by one: dp.shard_range)
    l->dp_global_batch = l->cfg.batch_size * (size_t)world_size;
    if (world_size > 1) {
        Player* p0 = l->players[0];
        double* d = nullptr;
        FI_CUDA_OK(cudaMalloc((void**)&d, sizeof(double)));
        const double mine = (double)l->cfg.batch_size;
        FI_CUDA_OK(cudaMemcpyAsync(d, &mine, sizeof(double), cudaMemcpyHostToDevice, p0->stream));
        FI_NCCL_OK(nccl().AllReduce(d, d, 1, ncclDouble, ncclSum, (ncclComm_t)p0->nccl_comm, p0->stream));
        double total = 0.0;
        FI_CUDA_OK(cudaMemcpyAsync(&total, d, sizeof(double), cudaMemcpyDeviceToHost, p0->stream));
        FI_CUDA_OK(cudaStreamSynchronize(p0->stream));
        cudaFree(d);
        l->dp_global_batch = (size_t)(total + 0.5);
    }
    return FI_OK;
}
int fi_learner_dp_world(const fi_learner* l) { return l ? l->dp_world : 0; }

// ---- inspection (parity tests) ---------------------------------------------------------------
__global__ void relu_mask_kernel(const float* __restrict__ a, const float* __restrict__ lo, size_t n,
                                 unsigned char* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (a[i] > 0.f || (lo && lo[i] > 0.f)) ? 1 : 0;
}

__global__ void relu_bits_expand_kernel(const uint32_t* __restrict__ bits, size_t n, unsigned char* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (bits[i >> 5] >> (i & 31)) & 1u;   // rows of 512 columns = 16 words: element i lives in word i / 32
}

int fi_learner_debug_relu_masks(fi_learner* l, int player, unsigned char* host, size_t n) {
    Player* p = get_player(l, player);
    if (!p || !host) return set_error(FI_ERR_ARG, "fi_learner_debug_relu_masks: null argument");
    const bool farmer = l->cfg.model == FI_MODEL_FARMER_LSTM;
    const size_t rows = farmer ? p->last_rows / l->cfg.entry_size : p->last_rows;
    const size_t per_layer = rows * fi::kHid;
    if (rows == 0 || n != 5 * per_layer)
        return set_error(FI_ERR_ARG, "fi_learner_debug_relu_masks: need n = 5 * %zu (rows of the last step x 512)", per_layer);
    FI_CUDA_OK(cudaSetDevice(l->cfg.device));
    std::lock_guard<std::mutex> step_lock(p->step_mu);
    unsigned char* dev = nullptr;
    FI_CUDA_OK(cudaMalloc((void**)&dev, n));
    int rc = FI_OK;
    for (int layer = 0; layer < 5 && rc == FI_OK; layer++) {
        if (const uint32_t* bits = farmer ? nullptr : fi::ac_relu_bits(p, layer)) {
            // the tensor-core path keeps the decisions its backward pass used as bit masks
            fi::LaunchScope ls("relu_bits_expand_kernel", p->stream, 1.125 * per_layer, fi::kWorkBytes);
            relu_bits_expand_kernel<<<fi::kNumSMs * 4, 256, 0, p->stream>>>(bits, per_layer, dev + layer * per_layer);
            rc = ls.done();
            continue;
        }
        const float *a = nullptr, *lo = nullptr;
        rc = farmer ? fi::farmer_activation(p, layer, &a, &lo) : fi::ac_activation(p, layer, &a, &lo);
        if (rc != FI_OK) break;
        fi::LaunchScope ls("relu_mask_kernel", p->stream, 5.0 * per_layer, fi::kWorkBytes);
        relu_mask_kernel<<<fi::kNumSMs * 4, 256, 0, p->stream>>>(a, lo, per_layer, dev + layer * per_layer);
        rc = ls.done();
    }
    if (rc == FI_OK && cudaMemcpyAsync(host, dev, n, cudaMemcpyDeviceToHost, p->stream) != cudaSuccess) rc = FI_ERR_CUDA;
    if (cudaStreamSynchronize(p->stream) != cudaSuccess && rc == FI_OK) rc = set_error(FI_ERR_CUDA, "stream sync failed");
    cudaFree(dev);
    return rc;
}

// ---- instrumentation ------------------------------------------------------------------------
void fi_prof_enable(int on) {
    fi::ProfState& p = fi::prof();
    std::lock_guard<std::mutex> g(p.mu);
    p.on.store(on != 0);
}

int fi_prof_collect(fi_prof_entry* out, int max_entries) {
    fi::ProfState& p = fi::prof();
    std::vector<fi::ProfRec> recs;
    {
        std::lock_guard<std::mutex> g(p.mu);
        recs.swap(p.recs);
    }
    std::vector<fi_prof_entry> agg;
    for (const fi::ProfRec& r : recs) {
        if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
        fi_prof_entry* e = nullptr;
        for (fi_prof_entry& a : agg)
            if (strncmp(a.name, r.name, sizeof(a.name)) == 0) e = &a;
        if (!e) {
            fi_prof_entry n;
            memset(&n, 0, sizeof(n));
            strncpy(n.name, r.name, sizeof(n.name) - 1);
            n.unit = r.unit;
            agg.push_back(n);
            e = &agg.back();
        }
        e->launches++;
        e->total_ms += ms;
        e->work += r.work;
    }
    cudaGetLastError();
    {
        std::lock_guard<std::mutex> g(p.mu);
        if (p.recs.empty()) p.used = 0;  // the event pool is reused by the next window
    }
    int n = 0;
    for (const fi_prof_entry& a : agg)
        if (out && n < max_entries) out[n++] = a;
    return (int)agg.size();
}

// ---- pinned host memory for callers that stage their own batches ------------------------------
void* fi_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        set_error(FI_ERR_CUDA, "fi_host_alloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
void fi_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
