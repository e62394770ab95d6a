#!/usr/bin/env python
"""Summarise an `ncu --set full` report (.ncu-rep) into a small JSON that can be committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_spmm.ncu-rep profiles/r01_spmm_ncu_full_summary.json

Reads the report with `ncu -i <rep> --page raw --csv` (works on the CPU box) and keeps, per profiled launch, the
metrics the roofline discussion in DESIGN.md / bench.py refers to: duration, DRAM bytes read/written (= `traffic`),
DRAM / L2 throughput, L2 hit rate, tensor-pipe activity, occupancy, registers, executed instructions."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.max",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "sm__inst_executed_pipe_lsu.sum",
    "l1tex__t_sector_hit_rate.pct",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        d = {"kernel": r[col["Kernel Name"]][:160]}
        for k in KEEP:
            if k in col:
                d[k] = {"value": r[col[k]], "unit": units[col[k]]}
        if "dram__bytes_read.sum" in d and "dram__bytes_write.sum" in d:
            def to_bytes(x):
                v, u = float(x["value"].replace(",", "")), x["unit"].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            d["traffic_bytes"] = to_bytes(d["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"])
        res.append(d)
    json.dump({"report": rep, "launches": res}, open(out, "w"), indent=1)
    for d in res:
        print(d["kernel"][:70], d.get("gpu__time_duration.sum", {}).get("value"), "traffic MB",
              round(d.get("traffic_bytes", 0) / 1e6, 1))


if __name__ == "__main__":
    main()
