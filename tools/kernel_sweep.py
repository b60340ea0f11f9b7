#!/usr/bin/env python
"""BASELINE.json configs[4]: gather / V-trace scan / fused Adam kernel sweep against the HBM roofline.

For every (batch M, length T) in {32..4096} x {20, 100, 400}: the ring gather (slot = T x 1024 B, algorithmic
2*M*slot bytes), the V-trace scan (24 B/transition + 4 B/trajectory) and, for several parameter counts, the fused
Adam update (28 B/param). Every kernel is launched through the C ABI's operator layer on the current torch
stream and timed with CUDA events; each timed launch works on a DIFFERENT buffer set out of a rotation whose
total footprint exceeds the 126 MB L2, so no launch re-reads lines a previous one left in L2.
CPU comparators (rank-0 host cores): the compiled reference SharedBuffer::readBatch when oracle/_ref exists (else
the oracle's C ring) and the oracle's float64 V-trace loop.

    python tools/kernel_sweep.py [--quick] > sweep.json
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

L2_BYTES = 126 * 2 ** 20


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def time_rotating(launch, nsets, iters):
    """launch(i, stream) enqueues one launch on buffer set i % nsets. The `iters` launches are captured into
    one CUDA graph (so the host's ~5 us per ctypes call does not bound the small kernels) and the graph is timed
    with CUDA events. Returns mean microseconds per launch."""
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(min(nsets, 3)):
            launch(i, side.cuda_stream)
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for i in range(iters):
                launch(i, torch.cuda.current_stream().cuda_stream)
        g.replay()
        side.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        g.replay()
        e1.record(side)
        side.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def nsets_for(bytes_per_set, cap_bytes=6 * 2 ** 30):
    n = max(2, int(np.ceil(1.5 * L2_BYTES / max(bytes_per_set, 1))) + 1)
    return int(min(n, max(2, cap_bytes // max(bytes_per_set, 1))))


def sweep_gather(M, T, peak):
    slot = T * 1024
    cap = M + 3
    per_set = (cap + M) * slot
    ns = nsets_for(per_set)
    rings = [torch.randint(0, 255, (cap * slot,), dtype=torch.uint8, device="cuda") for _ in range(ns)]
    outs = [torch.empty(M * slot, dtype=torch.uint8, device="cuda") for _ in range(ns)]
    def launch(i, s):
        j = i % ns
        fi.ops.gather(rings[j].data_ptr(), cap, slot, 2, M, outs[j].data_ptr(), s)  # first = 2: wraps past the ring end

    iters = max(20, min(400, int(2e9 // (2 * M * slot))))
    us = time_rotating(launch, ns, iters)
    algo = 2.0 * M * slot
    return {"kernel": "gather_slots_kernel", "M": M, "T": T, "bytes": algo, "us": us, "gbs": algo / us / 1e3,
            "frac": algo / us / 1e3 / peak, "buffer_sets": ns}


def sweep_vtrace(M, T, peak):
    n = M * T
    per_set = 24 * n + 4 * M
    ns = nsets_for(per_set)
    sets = []
    for _ in range(ns):
        sets.append([torch.randn(n, device="cuda") * 0.3, torch.full((n,), 0.99, device="cuda"), torch.randn(n, device="cuda"),
                     torch.randn(n, device="cuda"), torch.randn(M, device="cuda"), torch.empty(n, device="cuda"),
                     torch.empty(n, device="cuda")])
    def launch(i, s):
        a = sets[i % ns]
        fi.ops.vtrace(M, T, a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(), a[3].data_ptr(), a[4].data_ptr(),
                      a[5].data_ptr(), a[6].data_ptr(), stream=s)

    iters = max(20, min(400, int(1e9 // per_set)))
    us = time_rotating(launch, ns, iters)
    return {"kernel": "vtrace_scan_kernel", "M": M, "T": T, "bytes": per_set, "us": us, "gbs": per_set / us / 1e3,
            "frac": per_set / us / 1e3 / peak, "buffer_sets": ns}


def sweep_adam(n, peak):
    per_set = 28 * n
    ns = nsets_for(per_set)
    sets = [[torch.randn(n, device="cuda"), torch.randn(n, device="cuda"), torch.zeros(n, device="cuda"),
             torch.zeros(n, device="cuda")] for _ in range(ns)]
    step = [0]

    def launch(i, s):
        a = sets[i % ns]
        step[0] += 1
        fi.ops.adam("adam", 5e-4, step[0], n, a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(), a[3].data_ptr(), stream=s)

    iters = max(20, min(400, int(2e9 // per_set)))
    us = time_rotating(launch, ns, iters)
    return {"kernel": "fused_opt_kernel", "params": n, "bytes": per_set, "us": us, "gbs": per_set / us / 1e3,
            "frac": per_set / us / 1e3 / peak, "buffer_sets": ns}


def cpu_comparators(M, T):
    out = {}
    try:
        from oracle import pyoracle as po
        if os.path.exists(po.REF_HOST_SO):
            h = po.RefHost()
            sec = h.bench_read_batch(T, M + 1, M, 3)
            out["readBatch_reference"] = {"gbs": 3 * 2.0 * M * T * 1024 / sec / 1e9, "kind": "reference", "cores": 1,
                                          "what": "SharedBuffer::readBatch of the compiled reference, 3 batches"}
        o = po.Oracle()
        rng = np.random.default_rng(0)
        a = [rng.standard_normal((M, T)) for _ in range(4)] + [rng.standard_normal(M)]
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 0.5:
            o.vtrace(*a)
            reps += 1
        out["vtrace_oracle_f64"] = {"transitions_per_s": reps * M * T / (time.perf_counter() - t0), "kind": "port", "cores": 1}
    except Exception as e:  # comparators are optional
        out["error"] = str(e)[:200]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    peak = peaks()
    Ms = [32, 256, 1024, 4096] if args.quick else [32, 64, 128, 256, 512, 1024, 2048, 4096]
    Ts = [20, 100, 400]
    res = {"hbm_peak_gbs": peak, "device": torch.cuda.get_device_name(0), "gather": [], "vtrace": [], "adam": [], "cpu": {}}
    for M in Ms:
        for T in Ts:
            res["gather"].append(sweep_gather(M, T, peak))
            res["vtrace"].append(sweep_vtrace(M, T, peak))
            torch.cuda.empty_cache()
    # beyond configs[4]'s grid: the sizes at which the scan's working set leaves L2 and the launch ramp stops mattering
    # ("quote the size at which >= 70 % is reached")
    for M, T in [(8192, 400), (16384, 400)]:
        res["vtrace"].append(sweep_vtrace(M, T, peak))
        torch.cuda.empty_cache()
    for n in [1142801, 1514497, 8 * 2 ** 20, 64 * 2 ** 20, 256 * 2 ** 20]:
        res["adam"].append(sweep_adam(n, peak))
        torch.cuda.empty_cache()
    res["cpu"]["M1024_T100"] = cpu_comparators(1024, 100)
    print(json.dumps(res))
    for k in ("gather", "vtrace", "adam"):
        for r in res[k]:
            tag = f"M={r['M']:5d} T={r['T']:4d}" if "M" in r else f"n={r['params']:10d}"
            print(f"# {r['kernel']:22s} {tag} {r['bytes'] / 1e6:10.2f} MB {r['us']:9.2f} us {r['gbs']:8.1f} GB/s {r['frac']:.3f}", file=sys.stderr)


if __name__ == "__main__":
    main()
