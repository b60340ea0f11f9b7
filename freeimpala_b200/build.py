"""Build the sm_100a CUDA library in-tree: freeimpala_b200/_build/libfreeimpala_b200.so.

    python -m freeimpala_b200.build [--force]

nvcc cross-compiles without a GPU. The .so is git-ignored but travels to the GPU box with
the gpurun snapshot. One translation unit per nvcc process, compiled in parallel.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OUT_DIR = os.path.join(PKG, "_build")
LIB = os.path.join(OUT_DIR, "libfreeimpala_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function", "-Xptxas", "-v"]
# diagnostics build: the clock64 trace hooks of the tcgen05 GEMM and the recurrent kernels (FI_TC_TRACE / FI_LSTM_TRACE) are compiled
# in only on request -- `FI_TRACE_BUILD=1 python -m freeimpala_b200.build --force` (csrc/fi_internal.cuh says why)
if os.environ.get("FI_TRACE_BUILD", "0") == "1":
    NVCC_FLAGS.append("-DFI_TRACE_BUILD=1")


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps() -> list[str]:
    return (sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h")) +
            glob.glob(os.path.join(PKG, "host", "*")))


def up_to_date() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(os.path.join(OUT_DIR, "freeimpala_gpu")):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(f) <= t for f in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdr_time = max(os.path.getmtime(f) for f in _deps() if not f.endswith(".cu"))

    def compile_one(src: str) -> str:
        obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_time):
            return obj
        r = subprocess.run([nvcc, *NVCC_FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
        with open(obj[:-2] + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, sources()))
    r = subprocess.run([nvcc, "-shared", "-o", LIB + ".tmp", *objs, "-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(LIB + ".tmp", LIB)
    build_host()
    return LIB


HOST_BIN = os.path.join(OUT_DIR, "freeimpala_gpu")


def build_host() -> str:
    """The C++ host harness above the C ABI (freeimpala_b200/host): plain g++, linked against the library."""
    src = os.path.join(PKG, "host", "freeimpala_gpu_main.cpp")
    cxx = next((c for c in ("/usr/bin/g++", "g++") if not os.path.isabs(c) or os.path.exists(c)), "g++")
    r = subprocess.run([cxx, "-O2", "-std=c++17", "-pthread", "-Wall", src, "-o", HOST_BIN, "-L" + OUT_DIR,
                        "-lfreeimpala_b200", "-Wl,-rpath,$ORIGIN"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"host harness build failed:\n{r.stdout}\n{r.stderr}")
    return HOST_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
