"""SASS opcode histogram per kernel of the built library (CPU only: cuobjdump on the objects in freeimpala_b200/_build).

    python tools/sass_histogram.py > profiles/r2_sass_opcodes.md

What to look for (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor loads /
stores, UTCBAR = tcgen05.commit, SYNCS = mbarrier, ELECT = elect.sync; HMMA would be the legacy mma.sync path (there is none).
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INTEREST = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "ELECT", "R2UR", "HMMA",
            "FFMA", "FMUL", "FADD", "SHFL", "LDG", "STG", "LDS", "STS", "ATOMG", "REDG", "RED", "MUFU", "F2FP", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "freeimpala_b200", "_build", "*.o")))
    if not objs:
        sys.exit("build the library first: python -m freeimpala_b200.build")
    rows = []
    for obj in objs:
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        name, hist, total = None, None, 0
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                if name:
                    rows.append((os.path.basename(obj), name, total, hist))
                name, hist, total = m.group(1), collections.Counter(), 0
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
            if m and name:
                op = m.group(1)
                total += 1
                base = op.split(".")[0]
                hist[base] += 1
                if op.startswith("UTCHMMA.2CTA"):
                    hist["UTCHMMA.2CTA"] += 1
        if name:
            rows.append((os.path.basename(obj), name, total, hist))
    names = demangle([r[1] for r in rows])
    print("# SASS opcode histogram of libfreeimpala_b200.so (sm_100a), per kernel\n")
    print("`python tools/sass_histogram.py` (cuobjdump -sass on freeimpala_b200/_build/*.o). Counts are static instructions.\n")
    cols = [c for c in INTEREST if any(r[3].get(c) for r in rows)]
    print("| object | kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---|---|" + "---|" * len(cols))
    for obj, name, total, hist in rows:
        short = names.get(name, name)
        short = re.sub(r"\(.*", "", short).replace("fi::", "")
        short = re.sub(r"\bvoid\s+", "", short)
        print(f"| {obj} | `{short}` | {total} | " + " | ".join(str(hist.get(c, "")) if hist.get(c) else "" for c in cols) + " |")
    tc = [r for r in rows if r[3].get("UTCHMMA")]
    print(f"\n{len(tc)} kernel variants issue tcgen05.mma (UTCHMMA), {sum(1 for r in tc if r[3].get('UTCHMMA.2CTA'))} of them as CTA pairs "
          f"(UTCHMMA.2CTA); all of them read their accumulators with LDTM and load operands with UTMALDG. "
          f"No kernel contains HMMA (legacy mma.sync): {not any(r[3].get('HMMA') for r in rows)}.")


if __name__ == "__main__":
    main()
