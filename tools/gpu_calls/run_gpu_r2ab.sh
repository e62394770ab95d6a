#!/bin/bash
# round-2 call AB: L2 residency hints on the streaming SpMM (evict_last gathers, streaming row stores): A/B
mkdir -p gpurun_out
for m in 0 1; do
LLP_TUNING=5=$m timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fp32 --no-student --allow-tuning > gpurun_out/bench_hint$m.log 2>&1
python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/bench_hint$m.log") if x.startswith("{")][-1])
print("hint $m: collab value %.0f ms %.3f spmm frac %.3f share %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["share_of_step"]))
PY
done
LLP_TUNING=5=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:spmm_stream -c 4 --csv --log-file gpurun_out/ncu_hint.csv python tools/spmm_only.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/ncu_hint.csv') if not l.startswith('=='))]
hdr=rows[0]
for r in rows[1:]:
    d=dict(zip(hdr,r)); print(d['ID'], d['Kernel Name'][:50], d['Metric Name'], d['Metric Value'])
PY
LLP_TUNING=5=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "spmm" --timeout=400 --timeout-method=thread 2>&1 | tail -2
