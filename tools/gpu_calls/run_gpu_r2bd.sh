#!/bin/bash
# round-2 call BD: edge scorer producer fetching whole rows per warp: tests, micro-benchmark, A/B of the step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -m gpu -q -x --timeout=600 -k "edge or hadamard or scorer or plan or golden or captured" > gpurun_out/t_edge.log 2>&1; echo "edge tests exit=$?"; tail -n 2 gpurun_out/t_edge.log | cut -c1-300
timeout 300 python tools/kbench.py edgemlp 2>&1 | grep -i "edge\|MMA issue" | cut -c1-250
L=linkless_link_prediction_b200/libllp_b200.so
cp $L /tmp/lib_keep.so
for rep in 1 2; do
for v in prev new; do
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
