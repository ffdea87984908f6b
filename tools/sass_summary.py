#!/usr/bin/env python
"""Per-kernel SASS evidence that the hot path is Blackwell-native: counts of the tcgen05 / TMEM / TMA machine
instructions (profiling guide: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UTMAPF,
tcgen05.commit -> UTCBAR) in every kernel of the in-tree library, plus the legacy tensor-path mnemonics that must NOT
appear (HMMA = mma.sync / wmma).  Runs without a GPU:  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "xlstm_yolo_clean_b200", "lib", "libmlstm_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA", "MUFU", "SHFL", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in MNEMONICS:
                if op == k or op.startswith(k + ".") or (k in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM") and op.startswith(k)):
                    counts[cur][k] += 1
    names = demangle(order)
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS instruction counts per kernel (cuobjdump -sass, sm_100a)")
    print("# " + " ".join(f"{k:>8s}" for k in ["instrs"] + MNEMONICS) + "  kernel")
    tot = collections.Counter()
    for fn in order:
        c = counts[fn]
        tot.update(c)
        short = re.sub(r"mlstm::\(anonymous namespace\)::", "", names.get(fn, fn))
        short = re.sub(r"\(CUtensorMap.*", "(...)", short)
        short = re.sub(r"\(mlstm::.*", "(...)", short)
        print("  " + " ".join(f"{c[k]:8d}" for k in ["_total"] + MNEMONICS) + "  " + short[:110])
    print("# total")
    print("  " + " ".join(f"{tot[k]:8d}" for k in ["_total"] + MNEMONICS))
    assert tot["HMMA"] == 0, "legacy mma.sync / wmma instructions found"


if __name__ == "__main__":
    sys.exit(main())
