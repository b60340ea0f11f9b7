#!/usr/bin/env python
"""Where an end-to-end step goes on the GPU timeline (diagnostics for bench.py's `e2e`).

The e2e loop of bench.py with in-place producers (the ring + H2D + gather + step without the host memcpy) or copying writers,
with a CUDA event on the learner's stream before `readBatch` enqueues its wait-for-H2D + gather, one after it, and one after
`trainModel`: per step, how long the stream sat in wait + gather (the time the batch's last H2D copy was still in flight),
how long in the step, and the host-visible period.

    python tools/e2e_timeline.py [--steps 30] [--writers 14 | --inplace]
"""
import argparse
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--seq", type=int, default=100)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--writers", type=int, default=14)
    ap.add_argument("--inplace", action="store_true")
    ap.add_argument("--ring-batches", type=int, default=4)
    a = ap.parse_args()
    import torch
    import freeimpala_b200 as fi
    import bench
    M, T, K = a.batch, a.seq, a.steps
    slot_bytes = T * 1024
    L = fi.Learner(1, a.ring_batches * M, T, M, model="mlp_actor_critic", device=0, gemm_mode="auto", seed=1, lr=5e-4)
    lib = fi.load_library()
    stream_ptr = lib.fi_learner_stream(L._h, 0)
    ext = torch.cuda.ExternalStream(stream_ptr, device=torch.device("cuda", 0))
    host_ptr = lib.fi_host_alloc(2 * M * slot_bytes)
    host = np.ctypeslib.as_array((C.c_uint8 * (2 * M * slot_bytes)).from_address(host_ptr)).reshape(2 * M, slot_bytes)
    host[:M] = bench.synth_slots(1000, M, T)
    host[M:] = bench.synth_slots(2000, M, T)
    ring = L.getSharedBuffers()[0]

    def copy_writer(j, steps, nw):
        per = (M + nw - 1) // nw
        for s in range(steps):
            base = (s % 2) * M
            lo, hi = j * per, min(M, (j + 1) * per)
            for i in range(lo, hi, 32):
                ring.write_many(host[base + i:base + min(i + 32, hi)])

    def inplace_writer(j, steps, nw):
        for s in range(steps):
            for i in range(j * 64, M, nw * 64):
                n = min(64, M - i)
                ptrs, ticket = ring.reserve_many(n)
                C.c_uint32.from_address(ptrs[0] + 4 * 255).value = s
                ring.commit_many(ticket, n)

    def run(steps, writer, nw, record):
        ts = [threading.Thread(target=writer, args=(j, steps, nw)) for j in range(nw)]
        for t in ts:
            t.start()
        evs, host_t = [], []
        base = L.steps_done(0)
        for s in range(steps):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            h0 = time.perf_counter()
            e0.record(ext)
            b = ring.readBatch(M, stream_ptr)
            e1.record(ext)
            L.trainModel(0, b)
            e2.record(ext)
            if s > 0:
                L.losses_at(0, base + s)
            host_t.append(time.perf_counter() - h0)
            evs.append((e0, e1, e2))
        L.losses_at(0, base + steps)
        for t in ts:
            t.join()
        L.sync(0)
        if record:
            wait_gather = [a_.elapsed_time(b_) for a_, b_, _ in evs]
            step = [b_.elapsed_time(c_) for _, b_, c_ in evs]
            period = [evs[i][0].elapsed_time(evs[i + 1][0]) for i in range(len(evs) - 1)]
            idle = [evs[i][2].elapsed_time(evs[i + 1][0]) for i in range(len(evs) - 1)]
            med = lambda v: float(np.median(v[3:]))
            print(f"{'in-place' if writer is inplace_writer else 'copying'} producers x{nw}: stream period {med(period):.3f} ms = wait+gather "
                  f"{med(wait_gather):.3f} + step {med(step):.3f} + idle before the next readBatch reaches the stream {med(idle):.3f}; "
                  f"host loop {med(host_t) * 1e3:.3f} ms per step")

    if a.inplace:
        run(6, copy_writer, a.writers, False)   # leaves valid trajectories in the pinned slots
        run(K, inplace_writer, 2, True)
    else:
        run(6, copy_writer, a.writers, False)
        run(K, copy_writer, a.writers, True)
    L.close()
    lib.fi_host_free(host_ptr)


if __name__ == "__main__":
    main()
