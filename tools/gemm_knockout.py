"""Knock-out timing of the tcgen05 GEMM (GPU): re-runs bench.py with FI_TC_DBG bits set so that one part of the
kernel at a time does nothing (results are garbage; only the timings are read) and prints the average launch time
of the NT / NN / TN products. What disappears when a part is knocked out is what that part costs on the critical path.

    python tools/gemm_knockout.py [--steps 10] > gpurun_out/knockout.md
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BITS = {1: "no TMA loads", 2: "no MMAs", 4: "no promotion loads", 8: "no epilogue", 16: "no A_lo*B_hi MMA", 32: "no TMA stores"}
CASES = [0, 1, 2, 4, 8, 16, 32, 4 | 8, 1 | 4 | 8, 2 | 4 | 8, 1 | 2, 1 | 2 | 4, 1 | 2 | 4 | 8]


def main():
    steps = sys.argv[sys.argv.index("--steps") + 1] if "--steps" in sys.argv else "10"
    extra = []
    if "--batch" in sys.argv:
        extra = ["--batch", sys.argv[sys.argv.index("--batch") + 1]]
    names = ["NT", "NN", "TN", "NT,k162", "NT,head", "NN,head", "TN,head", "TN,n162"]
    print("| FI_TC_DBG | knocked out | ms/step | " + " | ".join(n + " us" for n in names) + " |")
    print("|---|---|---|" + "---|" * len(names))
    for flags in CASES:
        env = dict(os.environ, FI_TC_DBG=str(flags))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-e2e", "--no-cpu-baseline", "--steps", steps,
                            "--warmup", "3", *extra], env=env, capture_output=True, text=True)
        line = next((l for l in r.stdout.splitlines() if l.startswith("{")), None)
        what = ", ".join(v for b, v in BITS.items() if flags & b) or "-"
        if r.returncode != 0 or not line:
            print(f"| {flags} | {what} | failed rc={r.returncode} {r.stderr[-200:]!r} |")
            continue
        d = json.loads(line)
        k = d["kernels"]
        g = lambda n: k.get(f"gemm_tc_kernel<f16x3,{n}>", {}).get("avg_us", float("nan"))
        print(f"| {flags} | {what} | {d['ms_per_step']:.3f} | " + " | ".join(f"{g(n):.1f}" for n in names) + " |", flush=True)


if __name__ == "__main__":
    main()
