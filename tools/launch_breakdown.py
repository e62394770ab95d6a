#!/usr/bin/env python
"""Per-kernel breakdown of ONE training step from an ncu launch list (`--metrics gpu__time_duration.sum --csv`).

    python tools/launch_breakdown.py gpurun_out/launches.csv > profiles/rNN_step_breakdown.txt

A step starts at `llp::rng_advance_kernel` (one launch per step); the last complete step of the list is summarised.
ncu times are cold-cache and serialised: read the SHARES, not the absolute sum."""
import csv
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"'))]
    launches = [(r["Kernel Name"], float(r["Metric Value"]) / 1e3) for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
    starts = [i for i, (k, _) in enumerate(launches) if "rng_advance_kernel" in k]
    if len(starts) < 2:
        raise SystemExit("need at least two steps in the list")
    step = launches[starts[-2]:starts[-1]]
    agg = OrderedDict()
    for k, us in step:
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + us)
    total = sum(us for _, us in step)
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{c:3d} {t:8.1f} us {100 * t / total:5.1f}%  {k[:78]}")
    print(f"total {total:.1f} us, {len(step)} launches in one training step (ncu: cold-cache, serialised)")


if __name__ == "__main__":
    main()
