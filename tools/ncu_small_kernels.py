#!/usr/bin/env python
"""A fixed, short launch list of the HBM-bound kernels for `ncu --set full` (DRAM counters for the gather,
the V-trace scan, the fused loss head and the fused Adam): each kernel once at the learner's shape (batch 1024 x T=100) and once
at the size where the HBM roofline is approached (working set well beyond the 126 MB L2), every launch on fresh buffers.

    python tools/ncu_small_kernels.py                                  # plain run first (must exit 0)
    ncu --set full --clock-control none -k regex:"gather_slots|vtrace_scan|vtrace_loss_head|fused_opt" -o gpurun_out/r2_small \\
        python tools/ncu_small_kernels.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import freeimpala_b200 as fi


def main():
    s = torch.cuda.current_stream().cuda_stream
    for M, T in ((1024, 100), (4096, 100)):            # gather: 210 MB / 839 MB of traffic
        slot, cap = T * 1024, M + 3
        ring = torch.randint(0, 255, (cap * slot,), dtype=torch.uint8, device="cuda")
        out = torch.empty(M * slot, dtype=torch.uint8, device="cuda")
        fi.ops.gather(ring.data_ptr(), cap, slot, 2, M, out.data_ptr(), s)
        torch.cuda.synchronize()
        del ring, out
    for M, T in ((1024, 100), (4096, 400), (16384, 400)):   # scan: 2.5 MB / 39 MB / 157 MB
        n = M * T
        a = [torch.randn(n, device="cuda") * 0.3, torch.full((n,), 0.99, device="cuda"), torch.randn(n, device="cuda"),
             torch.randn(n, device="cuda"), torch.randn(M, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda")]
        fi.ops.vtrace(M, T, a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(), a[3].data_ptr(), a[4].data_ptr(), a[5].data_ptr(),
                      a[6].data_ptr(), stream=s)
        torch.cuda.synchronize()
        del a
    for M, T in ((1024, 100), (4096, 100)):            # fused loss head: reads the record fields out of the gathered batch
        rows = M * T
        batch = torch.randn(rows * 256, device="cuda")
        batch.view(rows, 256)[:, 178] = torch.randint(0, 16, (rows,), device="cuda").to(torch.int32).view(torch.float32)
        head = torch.randn(rows * 17, device="cuda")
        dhead = torch.empty(rows * 17, device="cuda")
        losses = torch.zeros(4, dtype=torch.float64, device="cuda")
        fi.ops.vtrace_loss_head(batch.data_ptr(), M, T, head.data_ptr(), 17, dhead.data_ptr(), losses.data_ptr(), stream=s)
        torch.cuda.synchronize()
        del batch, head, dhead
    for n in (1142801, 8 * 2 ** 20, 64 * 2 ** 20):     # fused Adam: 32 MB / 235 MB / 1.9 GB
        a = [torch.randn(n, device="cuda"), torch.randn(n, device="cuda"), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")]
        fi.ops.adam("adam", 5e-4, 1, n, a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(), a[3].data_ptr(), stream=s)
        torch.cuda.synchronize()
        del a
    print("ok")


if __name__ == "__main__":
    main()
