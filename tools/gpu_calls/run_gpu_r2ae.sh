#!/bin/bash
# round-2 call AE: 128-edge chunks with groups of 8 gathers (twice the bytes in flight per warp, same rounds per warp)
mkdir -p gpurun_out
timeout 600 python tools/kbench.py spmmab > gpurun_out/kbench_spmmab6.log 2>&1; grep "bfloat16" gpurun_out/kbench_spmmab6.log | cut -c1-200
for v in 0 3; do
LLP_TUNING=0=$v timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fp32 --no-student --allow-tuning > gpurun_out/bench_epw128_v$v.log 2>&1
python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/bench_epw128_v$v.log") if x.startswith("{")][-1])
print("kEPW=128 variant $v: collab value %.0f ms %.3f spmm frac %.3f share %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["share_of_step"]))
PY
done
