// Host-logic scenarios for fi_host.hpp (the reference-facing C++ host: SharedBuffer / ModelManager / Learner with the
// reference's method names) against the test double of the C ABI (fi_double.cpp). No GPU, no product library.
// Prints one JSON object; tests/test_host_logic.py asserts on it. Reference behaviour: include/freeimpala/learner.h:52-97
// (checkpointModel, workerThread), :158-197 (start / stop), data_structures.h:267-300 (readBatch + draining).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../freeimpala_b200/host/fi_host.hpp"

extern "C" {
void fi_double_fail_step_after(long n);
void fi_double_fail_read_after(long n);
long fi_double_read_calls();
size_t fi_double_consumed(fi_learner* l, int p, uint64_t* out, size_t max);
}

using namespace fi_host;
using clk = std::chrono::steady_clock;
static double since(clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); }

// (1) players x agents writing tagged trajectories; the workers consume them in per-writer FIFO order, versions advance by
// one per step, checkpoints appear every c iterations and at stop(), in the reference's file format.
static void scenario_run(const std::string& dir) {
    const size_t P = 2, B = 8, S = 2, M = 4, A = 6, T = 10, C = 3;
    const size_t iters = A * T / M;   // main.cpp:179
    Learner learner(P, B, S, M, 0, C, dir, "", iters);
    learner.start();
    std::vector<std::thread> actors;
    for (size_t a = 0; a < A; a++)
        actors.emplace_back([&, a] {
            auto bufs = learner.getSharedBuffers();
            std::vector<char> slot(S * ELEMENT_SIZE, (char)a);
            for (size_t it = 0; it < T; it++)
                for (size_t p = 0; p < P; p++) {
                    const uint64_t tag = (uint64_t)a << 32 | it;
                    memcpy(slot.data(), &tag, 8);
                    bufs[p]->write(slot);
                }
        });
    for (auto& t : actors) t.join();
    const auto t0 = clk::now();
    while ((learner.iterationsDone(0) < iters || learner.iterationsDone(1) < iters) && since(t0) < 20) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    learner.stop();
    bool fifo = true;
    size_t consumed = 0;
    for (size_t p = 0; p < P; p++) {
        std::vector<uint64_t> tags(A * T);
        const size_t n = fi_double_consumed(learner.handle(), (int)p, tags.data(), tags.size());
        consumed += n;
        std::vector<long> last(A, -1);
        for (size_t i = 0; i < n && i < tags.size(); i++) {
            const size_t a = tags[i] >> 32;
            const long it = (long)(tags[i] & 0xffffffffu);
            if (a >= A || it != last[a] + 1) fifo = false;   // every writer's trajectories arrive in the order it wrote them
            else last[a] = it;
        }
    }
    auto mm = learner.getModelManager();
    auto model = mm->getModel(0);
    FILE* f = fopen((dir + "/model_0_latest.bin").c_str(), "rb");
    uint64_t file_version = 0;
    long file_bytes = 0;
    if (f) {
        if (fread(&file_version, 8, 1, f) != 1) file_version = 0;
        fseek(f, 0, SEEK_END);
        file_bytes = ftell(f);
        fclose(f);
    }
    FILE* f3 = fopen((dir + "/model_1_3.bin").c_str(), "rb");   // the periodic checkpoint of iteration c = 3
    if (f3) fclose(f3);
    printf("\"run\": {\"iterations\": [%zu, %zu], \"expected\": %zu, \"versions\": [%llu, %llu], \"consumed\": %zu, \"fifo\": %s, "
           "\"model_version\": %llu, \"model_bytes\": %zu, \"latest_file_version\": %llu, \"latest_file_bytes\": %ld, "
           "\"periodic_checkpoint\": %s, \"updates_counted\": %llu}",
           learner.iterationsDone(0), learner.iterationsDone(1), iters, (unsigned long long)mm->getLatestVersion(0),
           (unsigned long long)mm->getLatestVersion(1), consumed, fifo ? "true" : "false", (unsigned long long)model->getVersion(),
           model->getData().size(), (unsigned long long)file_version, file_bytes, f3 ? "true" : "false",
           (unsigned long long)learner.stepMetrics().model_updates.load());
}

// (2) stop() with the workers blocked in readBatch on empty rings: draining wakes them (learner.h:170-172) and stop returns.
static void scenario_drain() {
    Learner learner(2, 8, 2, 4, 0, 0, "", "", 100);
    learner.start();
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
    auto bufs = learner.getSharedBuffers();
    std::vector<char> slot(2 * ELEMENT_SIZE, 1);
    bufs[0]->write(slot);   // fewer than a batch: readBatch must return the empty batch when draining
    const auto t0 = clk::now();
    learner.stop();
    printf("\"drain\": {\"stop_seconds\": %.3f, \"iterations\": %zu, \"left_in_ring\": %zu}", since(t0), learner.iterationsDone(0),
           bufs[0]->getFilledCount());
}

// (3) the step fails on its third call: the worker logs, stops and does not count the failed iteration; the other calls of
// the worker loop are not repeated (no spinning on a sticky failure).
static void scenario_step_failure() {
    fi_double_fail_step_after(2);
    Learner learner(1, 16, 1, 2, 0, 0, "", "", 100);
    learner.start();
    auto bufs = learner.getSharedBuffers();
    std::vector<char> slot(ELEMENT_SIZE, 2);
    for (int i = 0; i < 12; i++) bufs[0]->write(slot);
    const auto t0 = clk::now();
    while (learner.iterationsDone(0) < 2 && since(t0) < 10) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    std::this_thread::sleep_for(std::chrono::milliseconds(100));
    const long reads = fi_double_read_calls();
    std::this_thread::sleep_for(std::chrono::milliseconds(100));
    const long reads_later = fi_double_read_calls();
    learner.stop();
    fi_double_fail_step_after(-1);
    printf("\"step_failure\": {\"iterations\": %zu, \"updates_counted\": %llu, \"read_calls_while_idle\": %ld, \"failed\": %s}", learner.iterationsDone(0),
           (unsigned long long)learner.stepMetrics().model_updates.load(), reads_later - reads, learner.failed() ? "true" : "false");
}

// (4) readBatch itself fails (a sticky host-to-device failure in the product): the worker stops instead of retrying for ever.
static void scenario_read_failure() {
    const long before = fi_double_read_calls();
    fi_double_fail_read_after(1);
    Learner learner(1, 16, 1, 2, 0, 0, "", "", 100);
    learner.start();
    auto bufs = learner.getSharedBuffers();
    std::vector<char> slot(ELEMENT_SIZE, 3);
    for (int i = 0; i < 8; i++) bufs[0]->write(slot);
    std::this_thread::sleep_for(std::chrono::milliseconds(200));
    const long calls = fi_double_read_calls() - before;
    learner.stop();
    fi_double_fail_read_after(-1);
    printf("\"read_failure\": {\"iterations\": %zu, \"read_calls\": %ld}", learner.iterationsDone(0), calls);
}

// (5) SharedBuffer semantics through the shim: oversize write is refused, try_write refuses on a full ring, waitForModelUpdate
// times out without an update, a saved model loads back into a fresh learner with its version (-m / --starting-model).
static void scenario_api(const std::string& dir) {
    Learner a(1, 2, 1, 1, 0, 0, dir, "", 1);
    auto buf = a.getSharedBuffers()[0];
    std::vector<char> big(2 * ELEMENT_SIZE, 0), ok(ELEMENT_SIZE, 4);
    const bool oversize = buf->write(big);
    const bool w1 = buf->try_write(ok), w2 = buf->try_write(ok), w3 = buf->try_write(ok);
    const bool waited = a.getModelManager()->waitForModelUpdate(0, a.getModelManager()->getLatestVersion(0), 20);
    a.start();
    const auto t0 = clk::now();
    while (a.iterationsDone(0) < 1 && since(t0) < 10) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    a.stop();   // saves model_0_latest.bin with version 2
    Learner b(1, 2, 1, 1, 0, 0, "", dir, 1);
    printf("\"api\": {\"oversize_write\": %s, \"try_writes\": [%s, %s, %s], \"wait_without_update\": %s, \"resumed_version\": %llu, "
           "\"out_of_range_model\": %s}",
           oversize ? "true" : "false", w1 ? "true" : "false", w2 ? "true" : "false", w3 ? "true" : "false", waited ? "true" : "false",
           (unsigned long long)b.getModelManager()->getLatestVersion(0), b.getModelManager()->getModel(5) ? "true" : "false");
}

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : ".";
    printf("{");
    scenario_run(dir + "/run");
    printf(", ");
    scenario_drain();
    printf(", ");
    scenario_step_failure();
    printf(", ");
    scenario_read_failure();
    printf(", ");
    scenario_api(dir + "/api");
    printf("}\n");
    return 0;
}
