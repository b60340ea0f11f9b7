// TEST INFRASTRUCTURE ONLY (oracle/). Drop-in proof (VERDICT r1 missing #4): the reference's actor, include/freeimpala/agent.h,
// compiled UNMODIFIED from where it lies under /root/reference, against the fi_host classes through the alias headers in
// oracle/ref_shim/dropin/ -- then run the way cmd/freeimpala/main.cpp wires things (setupLearner :175-200, setupAgents
// :203-231, cleanup :234-260): a Learner with p players, `agents` Agent threads writing trajectories through
// SharedBuffer::write and pulling weights through ModelManager, the learner's worker threads stepping on the GPU.
// The reference's Agent fills its trajectories with rand() bytes (agent.h:62-72), so the losses mean nothing here; what is
// checked is the plumbing: every learner iteration runs, versions advance by one per update, agents see new versions, the
// reference's MetricsTracker counts the updates through the two calls fi_host::Learner::trainModel keeps (learner.h:34,48).
#include "freeimpala/learner.h"
#include FI_REF_AGENT_H   // "/root/reference/include/freeimpala/agent.h", unmodified

#include <cmath>
#include <cstdio>
#include <cstdlib>

int main(int argc, char** argv) {
    const size_t players = argc > 1 ? atoi(argv[1]) : 2, agents = argc > 2 ? atoi(argv[2]) : 2;
    const size_t S = 20, B = 8, M = 4, game_steps = 20, agent_iters = 8;
    const size_t T = agents * agent_iters / M;   // learner iterations: main.cpp:179
    spdlog::set_level(spdlog::level::warn);
    auto metrics = MetricsTracker::getInstance();
    metrics->start();
    std::string empty;
    Learner learner(players, B, S, M, /*train time*/ 0, /*checkpoint freq*/ 0, empty, empty, T);
    learner.start();
    std::vector<std::unique_ptr<Agent>> actors;
    std::vector<std::thread> threads;
    for (size_t a = 0; a < agents; a++)
        actors.push_back(std::make_unique<Agent>(a, players, S, game_steps, /*game time ms*/ 0, agent_iters, learner.getSharedBuffers(),
                                                 learner.getModelManager()));
    for (auto& a : actors) threads.emplace_back([&a] { a->run(); });
    for (auto& t : threads) t.join();
    // the workers finish their T iterations on what the agents wrote, then stop() drains
    for (int spin = 0; spin < 2000; spin++) {
        bool done = true;
        for (size_t p = 0; p < players; p++) done = done && learner.iterationsDone(p) >= T;
        if (done) break;
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
    learner.stop();
    bool ok = true;
    for (size_t p = 0; p < players; p++) {
        const uint64_t v = learner.getModelManager()->getLatestVersion(p);
        std::printf("player %zu: learner iterations %zu of %zu, published version %llu\n", p, learner.iterationsDone(p), T, (unsigned long long)v);
        ok = ok && learner.iterationsDone(p) == T && v == 1 + T;
    }
    // the reference's MetricsTracker exposes the update counter only as a rate: count = rate x elapsed
    const double secs = metrics->getTotalExecutionTime() / 1e9;
    const long updates = std::lround(metrics->getLearnerUpdatesPerSecond() * secs);
    const long syncs = std::lround(metrics->getAgentSyncsPerSecond() * secs);
    std::printf("reference MetricsTracker: learner model updates %ld (recordLearnerModelUpdate), training time %.3f ms (createTrainingTimer), "
                "agent iterations %llu, data transfers %llu, agent model syncs %ld\n",
                updates, metrics->getTotalTrainingTime() / 1e6, (unsigned long long)metrics->getTotalIterations(),
                (unsigned long long)metrics->getTotalDataTransfers(), syncs);
    ok = ok && updates == (long)(players * T) && metrics->getTotalTrainingTime() > 0 && metrics->getTotalIterations() == agents * agent_iters &&
         metrics->getTotalDataTransfers() == agents * agent_iters * players && syncs > 0;
    std::printf(ok ? "DROPIN_OK\n" : "DROPIN_FAILED\n");
    return ok ? 0 : 1;
}
