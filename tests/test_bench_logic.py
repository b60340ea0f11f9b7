"""Host-side logic of bench.py that needs no GPU: the per-kernel duration accounting of the roofline table."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


PEAKS = {"hbm_gbs": 6542.1, "bf16_tflops_sustained": 1400.7}


def _prof(K):
    # name: (launches per step, bracketed us per launch, work per launch, unit)
    rows = {"gemm": (12, 150.0, 53.7e9, "flops"), "gather": (1, 60.0, 209.7e6, "bytes"), "zero": (1, 6.0, 256.0, "bytes")}
    return {n: {"launches": lps * K, "total_ms": lps * K * us / 1e3, "work": lps * K * w, "unit": u} for n, (lps, us, w, u) in rows.items()}


def test_kernel_durations_add_up_to_the_headline_step(bench):
    K = 20
    prof = _prof(K)
    bracketed_step_ms = (12 * 150.0 + 60.0 + 6.0) / 1e3
    headline_ms = 1.60
    kt = bench.kernel_table(prof, K, bracketed_step_ms, 1, PEAKS, headline_ms)
    total_us = sum(v["avg_us"] * v["launches_per_step"] for v in kt.values())
    assert abs(total_us - headline_ms * 1e3) < 0.01
    c = kt["gemm"]["bracket_overhead_us"]
    assert 0 < c < 150.0 and kt["gather"]["bracket_overhead_us"] == pytest.approx(c)
    assert kt["zero"]["bracket_overhead_us"] == pytest.approx(3.0)          # never more than half of a bracket
    assert kt["gemm"]["avg_us_bracketed"] == pytest.approx(150.0) and kt["gemm"]["avg_us"] == pytest.approx(150.0 - c)
    assert kt["gemm"]["frac"] > kt["gemm"]["frac_bracketed"] == pytest.approx(53.7e9 / 150e-6 / 1e12 / PEAKS["bf16_tflops_sustained"])
    assert kt["gather"]["unit"] == "GB/s" and kt["gemm"]["unit"] == "TFLOP/s"


def test_no_correction_when_the_brackets_already_fit_or_with_several_ranks(bench):
    K = 10
    prof = _prof(K)
    bracketed_step_ms = (12 * 150.0 + 60.0 + 6.0) / 1e3
    for kt in (bench.kernel_table(prof, K, bracketed_step_ms, 1, PEAKS, 5.0),      # headline longer than the brackets
               bench.kernel_table(prof, K, bracketed_step_ms, 8, PEAKS, 1.6),      # only rank 0's brackets exist
               bench.kernel_table(prof, K, bracketed_step_ms, 1, PEAKS)):          # no headline given
        assert all(v["avg_us"] == v["avg_us_bracketed"] and v["bracket_overhead_us"] == 0.0 for v in kt.values())
    assert bench.kernel_table({}, K, 1.0, 1, PEAKS, 1.0) == {}
