// TEST DOUBLE of the C ABI (include/fi_learner.h) for the host-logic tests of fi_host.hpp -- test infrastructure only.
//
// tests/test_host_logic.py compiles this file together with host_logic_main.cpp so that fi_host::Learner's thread logic
// (worker loop, drain / stop, checkpoint threads, failure handling: reference include/freeimpala/learner.h:52-97, 158-197)
// can be exercised on a machine without a GPU. It is NOT part of the product and is never linked into
// libfreeimpala_b200.so: the product has no CPU path (fi_learner_create fails without a CUDA device). The double computes
// nothing: its "step" folds the batch bytes into a checksum blob and bumps the version, its ring is the reference's
// bounded FIFO (data_structures.h:219-300) over plain host memory, and it can be told to fail.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fi_learner.h"

struct fi_ring {
    size_t slot_bytes = 0, capacity = 0, read_index = 0, write_index = 0, count = 0;
    uint64_t consumed = 0;
    bool draining = false;
    std::vector<unsigned char> slots, batch;
    std::mutex mu;
    std::condition_variable not_full, not_empty;
};

struct fi_learner {
    fi_learner_config cfg;
    std::string ckpt;
    std::vector<fi_ring*> rings;
    std::vector<std::vector<unsigned char>> blob;   // per player: the "weights"
    std::vector<uint64_t> version, steps;
    std::vector<std::vector<uint64_t>> first_bytes;  // per player: first 8 bytes of every consumed slot, in consumption order
    std::mutex mu;
    std::condition_variable updated;
};

namespace {
thread_local std::string g_err;
std::atomic<long> g_fail_step_after{-1}, g_fail_read_after{-1}, g_steps{0}, g_reads{0}, g_read_calls{0};
constexpr size_t kBlob = 4096;
int fail(int code, const char* msg) { g_err = msg; return code; }
}  // namespace

extern "C" {
// knobs of the double (not part of the ABI)
void fi_double_fail_step_after(long n) { g_fail_step_after = n; g_steps = 0; }
void fi_double_fail_read_after(long n) { g_fail_read_after = n; g_reads = 0; }
long fi_double_read_calls() { return g_read_calls.load(); }
size_t fi_double_consumed(fi_learner* l, int p, uint64_t* out, size_t max) {
    std::lock_guard<std::mutex> g(l->mu);
    const size_t n = l->first_bytes[p].size() < max ? l->first_bytes[p].size() : max;
    for (size_t i = 0; i < n; i++) out[i] = l->first_bytes[p][i];
    return l->first_bytes[p].size();
}

const char* fi_last_error(void) { return g_err.c_str(); }

int fi_ring_write(fi_ring* r, const void* src, size_t n) {
    std::unique_lock<std::mutex> lock(r->mu);
    r->not_full.wait(lock, [&] { return r->count < r->capacity; });
    if (n > r->slot_bytes) return 0;
    memcpy(r->slots.data() + r->write_index * r->slot_bytes, src, n);
    r->write_index = (r->write_index + 1) % r->capacity;
    r->count++;
    r->not_empty.notify_one();
    return 1;
}
int fi_ring_try_write(fi_ring* r, const void* src, size_t n) {
    std::unique_lock<std::mutex> lock(r->mu, std::try_to_lock);
    if (!lock.owns_lock() || r->count >= r->capacity || n > r->slot_bytes) return 0;
    memcpy(r->slots.data() + r->write_index * r->slot_bytes, src, n);
    r->write_index = (r->write_index + 1) % r->capacity;
    r->count++;
    r->not_empty.notify_one();
    return 1;
}
int fi_ring_read_batch(fi_ring* r, size_t m, void* stream, fi_batch* out) {
    g_read_calls++;
    memset(out, 0, sizeof(*out));
    out->slot_bytes = r->slot_bytes;
    out->stream = stream;
    if (g_fail_read_after >= 0 && g_reads++ >= g_fail_read_after) return fail(FI_ERR_CUDA, "double: injected readBatch failure");
    if (m == 0 || m > r->capacity) return fail(FI_ERR_ARG, "double: bad batch size");
    std::unique_lock<std::mutex> lock(r->mu);
    r->not_empty.wait(lock, [&] { return r->count >= m || r->draining; });
    if (r->draining && r->count < m) return 0;
    r->batch.resize(m * r->slot_bytes);
    for (size_t i = 0; i < m; i++)
        memcpy(r->batch.data() + i * r->slot_bytes, r->slots.data() + ((r->read_index + i) % r->capacity) * r->slot_bytes, r->slot_bytes);
    r->read_index = (r->read_index + m) % r->capacity;
    r->count -= m;
    out->dev_ptr = r->batch.data();
    out->num_slots = m;
    out->seq = r->consumed;
    r->consumed += m;
    r->not_full.notify_all();
    return 1;
}
void fi_ring_set_draining(fi_ring* r) {
    { std::lock_guard<std::mutex> g(r->mu); r->draining = true; }
    r->not_empty.notify_all();
    r->not_full.notify_all();
}
size_t fi_ring_filled_count(fi_ring* r) { std::lock_guard<std::mutex> g(r->mu); return r->count; }
int fi_batch_to_host(const fi_batch* b, void* dst, size_t n) { memcpy(dst, b->dev_ptr, n); return FI_OK; }

void fi_learner_config_default(fi_learner_config* c) {
    memset(c, 0, sizeof(*c));
    c->num_players = 2; c->buffer_capacity = 20; c->entry_size = 100; c->batch_size = 5; c->publish_every = 1;
}
fi_learner* fi_learner_create(const fi_learner_config* c) {
    fi_learner* l = new fi_learner();
    l->cfg = *c;
    if (c->checkpoint_location) l->ckpt = c->checkpoint_location;
    for (int p = 0; p < c->num_players; p++) {
        fi_ring* r = new fi_ring();
        r->slot_bytes = c->entry_size * FI_ELEMENT_SIZE;
        r->capacity = c->buffer_capacity;
        r->slots.assign(r->slot_bytes * r->capacity, 0);
        l->rings.push_back(r);
        l->blob.emplace_back(kBlob, (unsigned char)p);
    }
    l->version.assign(c->num_players, 1);
    l->steps.assign(c->num_players, 0);
    l->first_bytes.resize(c->num_players);
    return l;
}
void fi_learner_destroy(fi_learner* l) {
    for (fi_ring* r : l->rings) delete r;
    delete l;
}
fi_ring* fi_learner_ring(fi_learner* l, int p) { return l->rings[p]; }
void* fi_learner_stream(fi_learner*, int) { return nullptr; }
int fi_learner_sync(fi_learner*, int) { return FI_OK; }
int fi_learner_step(fi_learner* l, int p, const fi_batch* b) {
    if (g_fail_step_after >= 0 && g_steps++ >= g_fail_step_after) return fail(FI_ERR_CUDA, "double: injected step failure");
    std::lock_guard<std::mutex> g(l->mu);
    const unsigned char* src = static_cast<const unsigned char*>(b->dev_ptr);
    for (size_t i = 0; i < b->num_slots; i++) {
        uint64_t w;
        memcpy(&w, src + i * b->slot_bytes, 8);
        l->first_bytes[p].push_back(w);
        for (size_t k = 0; k < kBlob; k++) l->blob[p][k] ^= src[i * b->slot_bytes + (k % b->slot_bytes)];
    }
    l->steps[p]++;
    l->version[p]++;
    l->updated.notify_all();
    return FI_OK;
}
int fi_learner_losses_at(fi_learner* l, int p, uint64_t step, float out[4]) {
    std::lock_guard<std::mutex> g(l->mu);
    if (step == 0 || step > l->steps[p]) return fail(FI_ERR_ARG, "double: no such step");
    out[0] = (float)step; out[1] = out[2] = out[3] = 0.f;
    return FI_OK;
}
size_t fi_model_bytes(const fi_learner*) { return kBlob; }
uint64_t fi_model_version(fi_learner* l, int p) { std::lock_guard<std::mutex> g(l->mu); return l->version[p]; }
int fi_model_get(fi_learner* l, int p, void* dst, size_t n, uint64_t* v) {
    std::lock_guard<std::mutex> g(l->mu);
    if (n != kBlob) return fail(FI_ERR_ARG, "double: blob size");
    memcpy(dst, l->blob[p].data(), n);
    if (v) *v = l->version[p];
    return FI_OK;
}
int fi_model_wait_update(fi_learner* l, int p, uint64_t cur, int timeout_ms) {
    std::unique_lock<std::mutex> lock(l->mu);
    return l->updated.wait_for(lock, std::chrono::milliseconds(timeout_ms), [&] { return l->version[p] > cur; }) ? 1 : 0;
}
int fi_model_save(fi_learner* l, int p, uint64_t it, int) {
    if (l->ckpt.empty()) return fail(FI_ERR_IO, "double: no checkpoint location");
    std::vector<unsigned char> blob;
    uint64_t v;
    { std::lock_guard<std::mutex> g(l->mu); blob = l->blob[p]; v = l->version[p]; }
    for (const std::string& name : {l->ckpt + "/model_" + std::to_string(p) + "_" + std::to_string(it) + ".bin",
                                    l->ckpt + "/model_" + std::to_string(p) + "_latest.bin"}) {
        FILE* f = fopen(name.c_str(), "wb");
        if (!f) return fail(FI_ERR_IO, "double: cannot open checkpoint file");
        fwrite(&v, 8, 1, f);                        // data_structures.h:105-110: u64 version + raw bytes
        fwrite(blob.data(), 1, blob.size(), f);
        fclose(f);
    }
    return FI_OK;
}
int fi_model_load(fi_learner* l, const char* dir) {
    for (size_t p = 0; p < l->blob.size(); p++) {
        FILE* f = fopen((std::string(dir) + "/model_" + std::to_string(p) + "_latest.bin").c_str(), "rb");
        if (!f) return fail(FI_ERR_IO, "double: no checkpoint");
        uint64_t v = 0;
        const bool ok = fread(&v, 8, 1, f) == 1 && fread(l->blob[p].data(), 1, kBlob, f) == kBlob;
        fclose(f);
        if (!ok) return fail(FI_ERR_IO, "double: short checkpoint");
        l->version[p] = v;
    }
    return FI_OK;
}
}  // extern "C"
