#!/bin/bash
# round-2 call U: hub table (small / big hubs, records) + zero-fill with one round trip: tests, fix-up time, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_config_sizes.py -m gpu -q -k "spmm or c4 or c5 or sage" --timeout=400 --timeout-method=thread > gpurun_out/t_spmm.log 2>&1; echo "spmm tests exit=$?"; tail -n 5 gpurun_out/t_spmm.log | cut -c1-200
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fp32 > gpurun_out/bench_collab_u.log 2>&1; echo "bench exit=$?"
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_collab_u.log") if x.startswith("{")][-1])
print("collab value %.0f ms %.3f e2e %.0f eval %.3f spmm frac %.3f share %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["eval"]["ms"], d["roofline"]["frac"], d["roofline"]["share_of_step"]))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spmm -c 16 --csv --log-file gpurun_out/ncu_fix.csv python tools/spmm_only.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/ncu_fix.csv') if not l.startswith('=='))]
hdr=rows[0]
for r in rows[1:]:
    d=dict(zip(hdr,r)); print(d['ID'], d['Kernel Name'][:60], d['Metric Value'])
PY
