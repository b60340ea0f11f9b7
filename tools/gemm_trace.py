"""(Needs a library with the trace hooks compiled in: FI_TRACE_BUILD=1 python -m freeimpala_b200.build --force.)
Pipeline trace of the tcgen05 GEMM (GPU): runs a few learner steps with FI_TC_TRACE set and prints, per role of CTA 0..3,
how long each pipeline event takes (median / p90 clocks between consecutive events of that role).

    FI_TC_TRACE=0,512,512 python tools/gemm_trace.py     # NT products with n >= 512 and k >= 512 (forward, big layers)

Tags: producer 1 stage free (after waiting for `empty`), 2 loads issued; issuer 10 TMEM buffer free (after `main_empty`),
11 stage full, 12 MMAs issued, 13 commit(empty) issued, 14 commit(main_full) issued; promotion warp 20 chunk ready
(after `main_full`), 21 chunk promoted + buffer released, 22 k-loop of the tile done (epilogue runs until the next 20).
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NAMES = {1: "P wait empty", 2: "P issue loads", 10: "I wait main_empty", 11: "I wait full", 12: "I issue MMAs", 13: "I commit empty",
         14: "I commit main_full", 20: "W wait main_full", 21: "W promote+release", 22: "W tile k-loop end (after last release)", 23: "W epi: scale, mask, relu, bits, colsum",
         24: "W epi: hi convert + amax", 25: "W epi: wait staging tile free (hi)", 26: "W epi: hi st.shared + TMA store",
         27: "W epi: lo' arithmetic", 28: "W epi: wait staging tile free (lo)", 29: "W epi: lo st.shared + TMA store"}


def main():
    os.environ.setdefault("FI_TC_TRACE", "0,512,512")
    os.environ["FI_GRAPH"] = "0"   # the trace buffer is (re)set with a legacy-stream memset at every traced launch
    import bench
    import freeimpala_b200 as fi
    import torch
    m = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 1024
    t = 100
    L = fi.Learner(1, m, t, m, model="mlp_actor_critic", seed=1)
    slots = bench.synth_slots(5, m, t)
    for _ in range(3):
        L.trainModel(0, L.stage_batch(0, slots))
    L.sync(0)
    n = 4 * 3 * 4096
    buf = np.zeros(n, np.uint64)
    got = fi.load_library().fi_debug_tc_trace(buf.ctypes.data, buf.nbytes)
    assert got == n, fi.last_error() if hasattr(fi, "last_error") else got
    buf = buf.reshape(4, 3, 4096)
    print(f"FI_TC_TRACE={os.environ['FI_TC_TRACE']} FI_TC_DBG={os.environ.get('FI_TC_DBG', '0')} batch {m}")
    for cta in range(2):
        for role, rname in enumerate(("producer", "issuer", "promotion warp 0")):
            ev = buf[cta, role]
            ev = ev[ev != 0]
            if ev.size < 8:
                continue
            tags, clk = (ev & 0xFF).astype(int), (ev >> 8).astype(np.int64)
            total = clk[-1] - clk[0]
            print(f"\nCTA {cta} {rname}: {ev.size} events over {total} clocks")
            d = np.diff(clk)
            for tag in sorted(set(tags[1:])):
                sel = d[tags[1:] == tag]
                sel_ss = sel[len(sel) // 8:] if len(sel) > 16 else sel   # steady state: skip the first eighth
                print(f"  -> {NAMES.get(tag, tag):40s} n={sel.size:5d} median {np.median(sel_ss):8.0f}  p90 {np.percentile(sel_ss, 90):8.0f}  "
                      f"sum {sel.sum():9d} ({100.0 * sel.sum() / total:5.1f} %)")
    # matched latencies of CTA 0: k-block i's loads issued (producer tag 2) -> seen full by the issuer (tag 11); stage released
    # by the issuer's commit (tag 13 of k-block i) -> seen free by the producer (tag 1 of k-block i + stages); chunk committed
    # (tag 14) -> seen by the promotion warp (tag 20); buffer released (tag 21) -> seen by the issuer (tag 10 two chunks later)
    def times(cta, role, tag):
        ev = buf[cta, role]
        ev = ev[ev != 0]
        return (ev[(ev & 0xFF) == tag] >> 8).astype(np.int64)
    def report(name, a, b):
        n = min(len(a), len(b))
        if n < 8:
            return
        d = (b[:n] - a[:n])[n // 8:]
        print(f"  {name:60s} median {np.median(d):8.0f}  p10 {np.percentile(d, 10):8.0f}  p90 {np.percentile(d, 90):8.0f}")
    print("\nmatched latencies (CTA 0):")
    issue, full = times(0, 0, 2), times(0, 1, 11)
    report("loads issued -> stage seen full by the issuer", issue, full)
    stages = int(os.environ.get("FI_TRACE_STAGES", "3"))
    commit, free = times(0, 1, 13), times(0, 0, 1)
    report(f"commit(empty) issued -> stage seen free by the producer ({stages} stages)", commit, free[stages:] if len(free) > stages else free)
    cfull, seen = times(0, 1, 14), times(0, 2, 20)
    report("commit(main_full) issued -> chunk seen by promotion warp 0", cfull, seen)
    rel, got = times(0, 2, 21), times(0, 1, 10)
    report("TMEM buffer released by warp 0 -> seen free by the issuer (2 chunks later)", rel, got[2:] if len(got) > 2 else got)
    # issuer timeline of CTA 0, a window in the steady state
    ev = buf[0, 1]
    ev = ev[ev != 0]
    tags, clk = (ev & 0xFF).astype(int), (ev >> 8).astype(np.int64)
    s = min(len(ev) // 2, 400)
    print("\nissuer timeline (CTA 0), clocks relative to the first shown event:")
    print(" ".join(f"{t_}@{c - clk[s]}" for t_, c in zip(tags[s:s + 40], clk[s:s + 40])))
    ev = buf[0, 2]
    ev = ev[ev != 0]
    tags2, clk2 = (ev & 0xFF).astype(int), (ev >> 8).astype(np.int64)
    w = np.searchsorted(clk2, clk[s])
    print("promotion warp 0 timeline, same origin:")
    print(" ".join(f"{t_}@{c - clk[s]}" for t_, c in zip(tags2[w:w + 24], clk2[w:w + 24])))
    ev = buf[0, 0]
    ev = ev[ev != 0]
    tags0, clk0 = (ev & 0xFF).astype(int), (ev >> 8).astype(np.int64)
    w = np.searchsorted(clk0, clk[s])
    print("producer timeline, same origin:")
    print(" ".join(f"{t_}@{c - clk[s]}" for t_, c in zip(tags0[w:w + 24], clk0[w:w + 24])))
    L.close()


if __name__ == "__main__":
    main()
