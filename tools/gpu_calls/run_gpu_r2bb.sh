#!/bin/bash
# round-2 call BB: A/B of the step with {old, new} edge plan x {old, new} edge scorer producer (four prebuilt libraries)
mkdir -p gpurun_out
L=linkless_link_prediction_b200/libllp_b200.so
cp $L /tmp/lib_keep.so
for rep in 1 2; do
for v in old plan scorer new; do
  cp tools/_build/lib_$v.so $L
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-fp32 --no-student > gpurun_out/bench_ab_$v.log 2>&1
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads([x for x in open(f"gpurun_out/bench_ab_{v}.log") if x.startswith("{")][-1])
print("%-7s ms %.4f e2e %.4f ms launches %s eval %.3f spmm share %.3f" % (v, d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("gpu_launches"), d["eval"]["ms"], d["roofline"]["share_of_step"]))
PY
done
done
cp /tmp/lib_keep.so $L
