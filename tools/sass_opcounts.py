#!/usr/bin/env python
"""Per-kernel SASS opcode counts of libllp_b200.so: which kernels are Blackwell-native (tcgen05 = UTC*MMA, TMEM loads =
LDTM, TMA = UTMALDG / UBLKCP) and which are CUDA-core kernels.  Runs anywhere (cuobjdump, no GPU):
    python tools/sass_opcounts.py > profiles/r02_sass_opcounts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "linkless_link_prediction_b200", "libllp_b200.so")
COLS = ["UTC*MMA", "UTC*MMA.2CTA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FFMA", "LDG", "STG", "LDS", "STS", "ATOM/RED"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = counts.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["total"] += 1
        if re.match(r"UTC[A-Z]*MMA", op):
            cur["UTC*MMA.2CTA" if ".2CTA" in op else "UTC*MMA"] += 1
        elif op.startswith("LDTM"):
            cur["LDTM"] += 1
        elif op.startswith("UTMALDG"):
            cur["UTMALDG"] += 1
        elif op.startswith("UBLKCP"):
            cur["UBLKCP"] += 1
        elif op.startswith("UTCBAR"):
            cur["UTCBAR"] += 1
        elif op.startswith("SYNCS"):
            cur["SYNCS"] += 1
        elif op.startswith("HMMA"):
            cur["HMMA"] += 1
        elif op.startswith("FFMA"):
            cur["FFMA"] += 1
        elif op.startswith("LDG"):
            cur["LDG"] += 1
        elif op.startswith("STG"):
            cur["STG"] += 1
        elif op.startswith("LDS"):
            cur["LDS"] += 1
        elif op.startswith("STS"):
            cur["STS"] += 1
        elif op.startswith("ATOM") or op.startswith("RED"):
            cur["ATOM/RED"] += 1
    names = demangle(list(counts))
    print(f"# SASS opcode counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; static instruction counts)")
    print(f"# tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG, cp.async.bulk -> UBLKCP, mma.sync -> HMMA")
    print(f"{'kernel':78s} {'instr':>6s} " + " ".join(f"{c:>8s}" for c in COLS))
    tc, other = [], []
    for k, c in counts.items():
        name = re.sub(r"\(.*", "", names.get(k, k))
        name = re.sub(r"^void ", "", name)
        row = f"{name[:78]:78s} {c['total']:6d} " + " ".join(f"{c[col]:8d}" for col in COLS)
        (tc if (c["UTC*MMA"] or c["UTC*MMA.2CTA"]) else other).append(row)
    print("## tensor-core kernels (tcgen05 / TMEM / TMA)")
    print("\n".join(tc))
    print("## CUDA-core kernels (HBM-bound byte / index work and small reductions)")
    print("\n".join(other))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print(f"## totals: kernels {len(counts)}, " + ", ".join(f"{col} {tot[col]}" for col in COLS))
    assert tot["HMMA"] == 0, "legacy mma.sync tensor path found"


if __name__ == "__main__":
    main()
