"""Shared helpers for the test-suite: deterministic synthetic weights / trajectories.

numpy's PCG64 stream is stable across numpy versions, so fixtures in tests/golden/ made by
tools/make_golden.py (which feeds these same arrays to the UNMODIFIED reference compiled in
oracle/_ref) can be re-derived on any box from the seeds alone.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# model.parameters() order of FarmerLstmModel (reference cmd/libtorch_bench/main.cpp:16-22)
FARMER_SHAPES = [(512, 162), (512, 128), (512,), (512,), (512, 612), (512,), (512, 512), (512,),
                 (512, 512), (512,), (512, 512), (512,), (512, 512), (512,), (1, 512), (1,)]
FARMER_FANIN = [128, 128, 128, 128, 612, 612, 512, 512, 512, 512, 512, 512, 512, 512, 512, 512]
# this build's MLP actor-critic: trunk shapes of main.cpp:17-21 + fused head [17,512]
AC_SHAPES = [(512, 162), (512,), (512, 512), (512,), (512, 512), (512,), (512, 512), (512,),
             (512, 512), (512,), (17, 512), (17,)]
AC_FANIN = [162, 162, 512, 512, 512, 512, 512, 512, 512, 512, 512, 512]


def init_params(shapes, fanin, seed: int) -> np.ndarray:
    """U(+-1/sqrt(fan_in)) like torch::nn defaults, flat fp32 in parameters() order."""
    rng = np.random.default_rng(seed)
    parts = []
    for shp, fi in zip(shapes, fanin):
        k = 1.0 / np.sqrt(fi)
        parts.append(rng.uniform(-k, k, size=int(np.prod(shp))).astype(np.float32))
    return np.concatenate(parts)


def farmer_params(seed: int) -> np.ndarray:
    return init_params(FARMER_SHAPES, FARMER_FANIN, seed)


def ac_params(seed: int) -> np.ndarray:
    return init_params(AC_SHAPES, AC_FANIN, seed)


def farmer_batch(seed: int, b: int, t: int):
    """z[b,t,162], x[b,484], target[b] ~ N(0,1) fp32 (make_batch, main.cpp:85-91)."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((b, t, 162)).astype(np.float32)
    x = rng.standard_normal((b, 484)).astype(np.float32)
    tg = rng.standard_normal((b,)).astype(np.float32)
    return z, x, tg


def vtrace_batch(seed: int, m: int, t: int, done_p: float = 0.01):
    """obs~N(0,1), behaviour logits~N(0,1), uniform actions,
    rewards~N(0,1), done~Bernoulli(p) -> discount 0.99(1-done), bootstrap~N(0,1)."""
    rng = np.random.default_rng(seed)
    obs = rng.standard_normal((m, t, 162)).astype(np.float32)
    mu = rng.standard_normal((m, t, 16)).astype(np.float32)
    act = rng.integers(0, 16, size=(m, t)).astype(np.int32)
    rew = rng.standard_normal((m, t)).astype(np.float32)
    disc = (0.99 * (rng.random((m, t)) >= done_p)).astype(np.float32)
    boot = rng.standard_normal((m,)).astype(np.float32)
    return obs, mu, act, rew, disc, boot


def rel_l2(a, b) -> float:
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def rel_max(a, b) -> float:
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def trimmed_rel_l2(a, b, drop_frac=0.005) -> float:
    """Relative L2 over all but the `drop_frac` worst elements (see AdamParity)."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    d = np.abs(a - b)
    keep = max(1, int(np.ceil(d.size * (1.0 - drop_frac))))
    idx = np.argpartition(d, keep - 1)[:keep]
    return float(np.linalg.norm(d[idx]) / max(np.linalg.norm(b[idx]), 1e-300))


class AdamParity:
    """Parameter parity after N optimiser steps, honest about Adam's conditioning.

    Adam's update is lr * m / (sqrt(v) + eps): on the first steps that is ~ lr * sign(g), so an element
    whose gradient is comparable to the fp32 noise of the computation can move by up to 2*lr in the
    opposite direction of the float64 oracle while every gradient agrees to 1e-6 (one such element
    among 1e6 already costs ~4e-5 of relative L2; libtorch's own fp32 step shows the same effect
    against float64). The north_star tolerance (relative error <= 1e-5 after N steps) is therefore
    asserted two ways:
      * teacher-forced: a float64 oracle that is fed the CUDA gradients each step must end with the
        same parameters, ALL elements, relative L2 <= 1e-5 (the gradients themselves are compared at
        1e-5 against the oracle's own at every step, at the same parameters);
      * free-running: an independent float64 oracle that never sees CUDA data. It cannot be held to 1e-5:
        besides the sign(g) effect, a ReLU unit whose pre-activation is within fp32 rounding of zero is
        decided differently by fp32 and float64 (a handful per step at 64 x 100), which moves whole
        weight-gradient rows of the layers below by ~1e-4 relative (measured; tools/diag_ac.py), and Adam
        turns that into parameter differences of the same order (measured after 3 steps at 64 x 100:
        6e-5 with the fp32 FFMA GEMMs, 1.1e-3 with the 3xTF32 tensor-core GEMMs, whose ~1e-6 forward
        error flips a few more units). It is kept as a sanity bound only: relative L2 <= 5e-3.
    """

    def __init__(self, forced, free):
        self.forced, self.free = forced, free

    def step(self, cuda_grads):
        """Call after both oracles ran loss_grad on this step's batch."""
        self.forced.set_grads(np.asarray(cuda_grads, np.float64))
        self.forced.opt_step()
        self.free.opt_step()

    def check(self, cuda_params, tol=1e-5):
        e_forced = rel_l2(cuda_params, self.forced.params())
        assert e_forced < tol, f"teacher-forced parameters differ: rel l2 {e_forced:.3e}"
        e_trim = trimmed_rel_l2(cuda_params, self.free.params())
        e_all = rel_l2(cuda_params, self.free.params())
        assert e_all < 5e-3, f"free-running parameters differ: rel l2 {e_all:.3e}"
        return e_forced, e_trim, e_all
