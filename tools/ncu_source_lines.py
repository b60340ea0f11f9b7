#!/usr/bin/env python
"""Top CUDA source lines of one captured launch by warp-stall samples and by executed warp instructions, read here
(no GPU) from an `ncu --set full --import-source on` report of a library built with -lineinfo.

    python tools/ncu_source_lines.py gpurun_out/x.ncu-rep <launch index> [top N] >> profiles/name.md
"""
import csv
import io
import subprocess
import sys


def main(rep, launch, top=14):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--launch-skip", str(launch),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    cur, kernel, data = None, None, []
    for r in csv.reader(io.StringIO(out)):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) == 2 and r[0] == "Function Name":
            kernel = kernel or r[1].split("(CUtensorMap")[0].replace("void ", "")
        elif len(r) > 8 and r[0] != "Line No" and r[2] == "-":   # a CUDA line row (its SASS rows follow, with addresses)
            try:
                data.append((int(r[6]), int(r[7]), cur, r[0], " ".join(r[1].split())[:110]))
            except ValueError:
                pass
    ts, ti = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
    print(f"\n### launch {launch}: `{kernel}`\n\n{ts} stall samples, {ti} warp instructions executed\n")
    print("| samples | instructions | line | source |\n|---|---|---|---|")
    for s, i, f, ln, src in sorted(data, reverse=True)[:top]:
        print(f"| {100 * s / ts:.1f} % | {100 * i / ti:.1f} % | `{f}:{ln}` | `{src.replace('|', '/')}` |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 14)
