#!/bin/bash
# round-2 call AD: final check of the committed tree: whole GPU suite, smoke(), default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --timeout-method=thread > gpurun_out/t_all.log 2>&1; echo "all exit=$?"; tail -n 3 gpurun_out/t_all.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1 ) 2>&1 | grep real; echo "bench default exit=$?"
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_default.log") if x.startswith("{")][-1])
print("collab value %.0f ms %.3f e2e %.0f eval %.3f spmm frac %.3f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["eval"]["ms"], d["roofline"]["frac"], d["gpu_launches"]))
print("fp32 ms", d["fp32"]["ms_per_step"], "student ms", d["student"]["ms_per_step"], "frac", d["student"]["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-400
