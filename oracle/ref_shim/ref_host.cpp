// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// C shim around the UNMODIFIED reference host data structures, compiled from where they
// lie under /root/reference/include (see oracle/Makefile). Exposes, for parity tests:
//   SharedBuffer   include/freeimpala/data_structures.h:191-307
//   Model          include/freeimpala/data_structures.h:43-157
//   ModelManager   include/freeimpala/data_structures.h:310-481
//   Learner        include/freeimpala/learner.h:7-208   (the stub trainModel loop)
// data_structures.h:141 uses std::optional without including it; the reference only builds
// because argparse.hpp pulls it in first, so we include <optional> ahead of it.
#include <optional>
#include "freeimpala/learner.h"

#include <chrono>
#include <cstdint>
#include <cstring>

extern "C" {

// ---- SharedBuffer -----------------------------------------------------------------
void* ref_ring_create(size_t entry_size, size_t capacity) { return new SharedBuffer(entry_size, capacity); }
void ref_ring_destroy(void* r) { delete static_cast<SharedBuffer*>(r); }
int ref_ring_write(void* r, const char* data, size_t n) {
    std::vector<char> v(data, data + n);
    return static_cast<SharedBuffer*>(r)->write(v) ? 1 : 0;
}
int ref_ring_try_write(void* r, const char* data, size_t n) {
    std::vector<char> v(data, data + n);
    return static_cast<SharedBuffer*>(r)->try_write(v) ? 1 : 0;
}
// Returns the number of entries in the batch (0 = empty batch on drain); entries are
// concatenated into out (caller provides batch_size * slot_bytes).
size_t ref_ring_read_batch(void* r, size_t batch_size, char* out) {
    auto batch = static_cast<SharedBuffer*>(r)->readBatch(batch_size);
    for (auto& e : batch) {
        std::memcpy(out, e.data(), e.size());
        out += e.size();
    }
    return batch.size();
}
void ref_ring_set_draining(void* r) { static_cast<SharedBuffer*>(r)->setDraining(); }
size_t ref_ring_filled_count(void* r) { return static_cast<SharedBuffer*>(r)->getFilledCount(); }

// Single-thread readBatch timing for the gather CPU baseline:
// fills the ring, reads `batch_size` slots, `iters` times; returns seconds in readBatch only.
double ref_ring_bench_read_batch(size_t entry_size, size_t capacity, size_t batch_size, int iters) {
    SharedBuffer ring(entry_size, capacity);
    std::vector<char> slot(entry_size * ELEMENT_SIZE);
    for (size_t i = 0; i < slot.size(); ++i) slot[i] = static_cast<char>(i * 131u + 7u);
    double total = 0;
    for (int it = 0; it < iters; ++it) {
        for (size_t i = 0; i < batch_size; ++i) ring.write(slot);
        auto t0 = std::chrono::steady_clock::now();
        auto batch = ring.readBatch(batch_size);
        auto t1 = std::chrono::steady_clock::now();
        total += std::chrono::duration<double>(t1 - t0).count();
        if (batch.size() != batch_size) return -1.0;
    }
    return total;
}

// ---- Model / ModelManager ----------------------------------------------------------
void* ref_mm_create(size_t num_players, size_t model_size, const char* dir) {
    return new ModelManager(num_players, model_size, dir);
}
void ref_mm_destroy(void* m) { delete static_cast<ModelManager*>(m); }
uint64_t ref_mm_latest_version(void* m, size_t p) { return static_cast<ModelManager*>(m)->getLatestVersion(p); }
// Publish `data` as the next version of player p, the way Learner::trainModel does
// (learner.h:40-45): createCopy -> new content, version+1 -> updateModel.
void ref_mm_publish(void* m, size_t p, const char* data, size_t n) {
    auto* mm = static_cast<ModelManager*>(m);
    auto next = mm->getModel(p)->createCopy();
    next->update(std::vector<char>(data, data + n));
    mm->updateModel(p, next);
}
uint64_t ref_mm_get(void* m, size_t p, char* out, size_t n) {
    auto model = static_cast<ModelManager*>(m)->getModel(p);
    auto d = model->getData();
    std::memcpy(out, d.data(), d.size() < n ? d.size() : n);
    return model->getVersion();
}
void ref_mm_save(void* m, size_t p, uint64_t iter) { static_cast<ModelManager*>(m)->saveModel(p, iter); }
void ref_mm_load(void* m, const char* dir) { static_cast<ModelManager*>(m)->loadModels(dir); }

// ---- Learner (stub trainModel) -----------------------------------------------------
// Runs the reference threaded learner loop with `writers` producer threads per player
// each writing `per_writer` trajectories; returns learner model updates completed and
// elapsed seconds. Used as the CPU comparator for the threaded config (BASELINE configs[2]).
int ref_learner_run(size_t players, size_t capacity, size_t entry_size, size_t batch_size,
                    size_t train_ms, size_t writers, size_t per_writer, const char* ckpt_dir,
                    double* seconds_out, uint64_t* updates_out) {
    size_t total = (writers * per_writer) / batch_size;
    auto metrics = MetricsTracker::getInstance();
    metrics->start();
    auto t0 = std::chrono::steady_clock::now();
    {
        Learner learner(players, capacity, entry_size, batch_size, train_ms, 0, ckpt_dir, "", total);
        learner.start();
        auto rings = learner.getSharedBuffers();
        std::vector<std::thread> th;
        for (size_t p = 0; p < players; ++p)
            for (size_t w = 0; w < writers; ++w)
                th.emplace_back([&, p, w] {
                    std::vector<char> slot(entry_size * ELEMENT_SIZE, static_cast<char>(w));
                    for (size_t i = 0; i < per_writer; ++i) rings[p]->write(slot);
                });
        for (auto& t : th) t.join();
        // wait for the learner workers to consume what was written
        auto mm = learner.getModelManager();
        for (size_t p = 0; p < players; ++p)
            while (rings[p]->getFilledCount() >= batch_size) std::this_thread::yield();
        auto t1 = std::chrono::steady_clock::now();
        *seconds_out = std::chrono::duration<double>(t1 - t0).count();
        uint64_t upd = 0;
        for (size_t p = 0; p < players; ++p) upd += mm->getLatestVersion(p) - 1;
        *updates_out = upd;
        learner.stop();
    }
    return 0;
}

}  // extern "C"
