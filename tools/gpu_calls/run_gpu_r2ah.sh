#!/bin/bash
# round-2 call AH: pair kernel (half warp per row, one 128-column slice per pass): parity tests, A/B timing, bench, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_config_sizes.py tests/test_gpu_e2e.py -m gpu -q --timeout=600 --timeout-method=thread > gpurun_out/t_pair.log 2>&1; echo "tests exit=$?"; tail -n 8 gpurun_out/t_pair.log | cut -c1-250
timeout 600 python tools/kbench.py spmmab > gpurun_out/kbench_spmmab7.log 2>&1; grep "bfloat16" gpurun_out/kbench_spmmab7.log | cut -c1-200
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fp32 --no-student > gpurun_out/bench_pair.log 2>&1; echo "bench exit=$?"
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_pair.log") if x.startswith("{")][-1])
print("pair kernel: collab value %.0f ms %.3f e2e %.0f eval %.3f spmm frac %.3f share %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["eval"]["ms"], d["roofline"]["frac"], d["roofline"]["share_of_step"]))
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:spmm_pair -c 4 --csv --log-file gpurun_out/ncu_pair.csv python tools/spmm_only.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/ncu_pair.csv') if not l.startswith('=='))]
hdr=rows[0]
for r in rows[1:]:
    d=dict(zip(hdr,r)); print(d['ID'], d['Kernel Name'][:40], d['Metric Name'], d['Metric Value'])
PY
