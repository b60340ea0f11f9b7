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


class AdamParity:
    """Parameter parity after N optimiser steps, honest about Adam's conditioning.

    Adam's update is lr * m / (sqrt(v) + eps): on the first steps that is ~ lr * sign(g), so an element
    whose gradient is smaller than the fp32 noise of the computation (|g| <~ 1e-6 * rms(g); a handful
    per million) can move by up to 2*lr in the opposite direction of the float64 oracle while every
    gradient agrees to 1e-6. One such element among 1e6 already costs ~4e-5 of relative L2. The
    north_star tolerance (relative error <= 1e-5) is therefore asserted over the elements whose
    gradient was well above the noise floor at EVERY step (|g| > tau * rms(g), tau = 1e-3: ~99.9 % of
    the non-dead parameters); the remaining elements must stay within the bounded update (2 * lr per
    step), and the relative L2 over ALL parameters is checked at 1e-3.
    """

    def __init__(self, lr: float, tau: float = 1e-3):
        self.lr, self.tau, self.mask, self.steps = lr, tau, None, 0

    def observe(self, oracle_grads):
        g = np.abs(np.asarray(oracle_grads, np.float64))
        nz = g[g > 0]
        rms = np.sqrt((nz ** 2).mean()) if nz.size else 0.0
        ok = (g > self.tau * rms) | (g == 0)        # exact zeros (dead units) stay exact on both sides
        self.mask = ok if self.mask is None else (self.mask & ok)
        self.steps += 1

    def check(self, got, want, tol=1e-5):
        got = np.asarray(got, np.float64)
        want = np.asarray(want, np.float64)
        m = self.mask
        assert m.mean() > 0.99, m.mean()
        err_ok = np.linalg.norm((got - want)[m]) / np.linalg.norm(want[m])
        assert err_ok < tol, f"well-conditioned parameters differ: rel l2 {err_ok:.3e}"
        assert np.abs(got - want)[~m].max(initial=0.0) <= 2.0 * self.lr * self.steps * 1.01
        assert rel_l2(got, want) < 1e-3
        return err_ok
