#!/usr/bin/env python
"""Host time to ENQUEUE one learner step (no synchronisation inside the measured region), against the device time
of the same steps. A step whose enqueue time approaches its device time is host-bound; with 8 ranks on a 16-core box
each rank has two cores for its Python thread, the publish thread and NCCL's proxy.

    python tools/host_enqueue.py [--batch 1024 --seq 100 --steps 4]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import freeimpala_b200 as fi
from freeimpala_b200._lib import FiBatch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--seq", type=int, default=100)
    ap.add_argument("--steps", type=int, default=4)   # few enough that the launch queue (about 1000 entries) never fills
    ap.add_argument("--model", default="mlp_actor_critic")
    a = ap.parse_args()
    M, T = a.batch, a.seq
    L = fi.Learner(1, 2 * M, T, M, model=a.model, seed=1)
    lib = fi.load_library()
    st = lib.fi_learner_stream(L._h, 0)
    ext = torch.cuda.ExternalStream(st)
    rng = np.random.default_rng(0)
    batch = torch.from_numpy(rng.standard_normal((M, T * 256)).astype(np.float32)).cuda()
    w = batch.view(M, T, 256)
    w[:, :, 178] = torch.from_numpy(rng.integers(0, 16, (M, T)).astype(np.int32)).cuda().view(torch.float32)
    w[:, :, 180] = 0.99
    raw = FiBatch(batch.data_ptr(), M, T * 1024, st, 0)
    b = fi.Batch(raw)
    out = {}
    for prof in (False, True):
        for _ in range(3):
            L.trainModel(0, b)
        L.sync(0)
        fi.prof_collect()
        fi.prof_enable(prof)
        n0 = fi.kernel_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per = []
        e0.record(ext)
        for _ in range(a.steps):
            t0 = time.perf_counter()
            L.trainModel(0, b)
            per.append((time.perf_counter() - t0) * 1e3)
        e1.record(ext)
        L.sync(0)
        fi.prof_enable(False)
        fi.prof_collect()
        out["prof_on" if prof else "prof_off"] = {
            "host_enqueue_ms_per_step": per, "device_ms_per_step": e0.elapsed_time(e1) / a.steps,
            "launches_per_step": (fi.kernel_launch_count() - n0) / a.steps}
    print(json.dumps(out))
    L.close()


if __name__ == "__main__":
    main()
