#!/usr/bin/env python
"""Informational comparator: the FarmerLstm training step in STOCK PyTorch on the same GPU.

This is what the reference's scripts/gpu_benchmark.py (--gpu cuda) times: torch.nn.LSTM (cuDNN) + torch.nn.Linear (cuBLAS) +
torch.optim.Adam, MSE loss, synthetic batch. The reference script itself cannot travel to the GPU box (/root/reference does not
exist there), so the model is restated here from the reference's C++ definition (cmd/libtorch_bench/main.cpp:14-42: LSTM
162 -> 128, batch_first; dense 612 -> 512 x5 with ReLU -> 1); nothing of this repository's CUDA code runs in this script. Timing
as bench.py does it: CUDA events around K steps after W warm-up steps, inputs resident on the device.

    python tools/torch_gpu_comparator.py [--batch 1024 --seq 100 --steps 20 --warmup 5]

Two lines: fp32 with TF32 off (the precision class this repository's step holds: parameters within 1e-5 of the reference) and
with TF32 on (torch's cuDNN default for RNNs; 10-bit mantissa products). Not a parity oracle and not a bench arm: one number a
maintainer would ask for, kept under profiles/.
"""
from __future__ import annotations

import argparse
import json

import torch
import torch.nn as nn


class Farmer(nn.Module):
    def __init__(self):
        super().__init__()
        self.lstm = nn.LSTM(162, 128, batch_first=True)
        widths = [128 + 484, 512, 512, 512, 512, 512]
        self.dense = nn.ModuleList(nn.Linear(a, b) for a, b in zip(widths[:-1], widths[1:]))
        self.out = nn.Linear(512, 1)

    def forward(self, z, x):
        h = self.lstm(z)[0][:, -1, :]
        a = torch.cat([h, x], dim=-1)
        for d in self.dense:
            a = torch.relu(d(a))
        return self.out(a)


def run(batch: int, seq: int, steps: int, warmup: int, tf32: bool) -> dict:
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = Farmer().to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    z = torch.randn(batch, seq, 162, device=dev)
    x = torch.randn(batch, 484, device=dev)
    y = torch.randn(batch, 1, device=dev)
    loss_fn = nn.MSELoss()

    def step():
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(model(z, x), y)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    return {"impl": "stock_pytorch_gpu", "torch": torch.__version__, "tf32": tf32, "ms_per_step": ms,
            "value": batch * seq / (ms * 1e-3), "unit": "transitions/s",
            "config": {"workload": f"FarmerLstm {batch}x{seq} MSE/Adam, cuDNN LSTM + cuBLAS, fp32{' (TF32 products)' if tf32 else ''}"},
            "steps": steps, "warmup": warmup, "gpu": torch.cuda.get_device_name(0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--seq", type=int, default=100)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    for tf32 in (False, True):
        print(json.dumps(run(args.batch, args.seq, args.steps, args.warmup, tf32)), flush=True)


if __name__ == "__main__":
    main()
