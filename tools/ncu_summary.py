#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, without a GPU): one row of key counters per captured launch,
the top warp-stall instructions of the first launch, and a traffic JSON that bench.py attaches to `roofline.traffic`.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/name   -> profiles/name.md, profiles/name_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB rd"), ("dram__bytes_write.sum", "MB wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]


def page(rep, name):
    return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout


def label(kname):
    if "gemm_tc_kernel" in kname:
        a = kname.split("<")[1].split(">")[0].replace("(int)", "").replace("(bool)", "").replace(" ", "").split(",")
        fmt = "f16x3," if len(a) > 4 and a[4] == "1" else ""
        pair = " pair" if len(a) > 3 and a[3] == "1" else ""
        return f"gemm_tc_kernel<{fmt}{'NT' if a[2] == '0' else ('NN' if a[1] == '0' else 'TN')}> BN={a[0]}{pair}"
    return kname.split("(")[0].replace("void ", "").replace("fi::", "")


def main(rep, out):
    rows = list(csv.reader(io.StringIO(page(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    lines = ["| launch | kernel | " + " | ".join(n for _, n in KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
    traffic = {}
    for n, r in enumerate(rows[2:]):
        vals = []
        for k, _ in KEYS:
            if k not in hdr:
                vals.append("-")
                continue
            v, u = float(r[hdr.index(k)].replace(",", "")), units[hdr.index(k)]
            if u == "Gbyte": v *= 1e3
            if u == "Kbyte": v /= 1e3
            if u == "byte": v /= 1e6
            if u == "ns": v /= 1e3
            if u == "ms": v *= 1e3
            vals.append(f"{v:.1f}" if v < 1e5 else f"{v:.0f}")
        lab = label(r[ki])
        lines.append(f"| {n} | `{lab}` | " + " | ".join(vals) + " |")
        t = traffic.setdefault(lab, {"launches": 0, "dram_mb": 0.0, "us": 0.0})
        t["launches"] += 1
        t["dram_mb"] += float(vals[1]) + float(vals[2])
        t["us"] += float(vals[0])
    for t in traffic.values():
        t["dram_bytes_per_launch"] = t["dram_mb"] * 1e6 / t["launches"]
        t["us_per_launch"] = t["us"] / t["launches"]
    src = page(rep, "source")
    block = src.split('"Kernel Name",')[1] if '"Kernel Name",' in src else ""
    stall = []
    if block:
        b = block.splitlines()
        rr = list(csv.reader(b[1:]))
        h = rr[0]
        if "Warp Stall Sampling (All Samples)" in h:
            si, so = h.index("Warp Stall Sampling (All Samples)"), h.index("Source")
            data = sorted(((int(x[si] or 0), x[so].strip()) for x in rr[1:] if len(x) > si), reverse=True)
            tot = sum(d[0] for d in data) or 1
            stall = [f"| {100 * s / tot:.1f} % | `{t[:100]}` |" for s, t in data[:12]]
    with open(out + ".md", "a") as f:
        f.write("\n".join(lines) + "\n")
        if stall:
            f.write("\nTop warp-stall samples, first captured launch (SASS):\n\n| share | instruction |\n|---|---|\n" + "\n".join(stall) + "\n")
    json.dump(traffic, open(out + "_traffic.json", "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
