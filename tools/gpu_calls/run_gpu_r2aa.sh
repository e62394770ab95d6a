#!/bin/bash
# round-2 call AA: default bench line with the appended student object; SpMM parity subset
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1 ) 2>&1 | grep real; echo "bench default exit=$?"
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_default.log") if x.startswith("{")][-1])
print("collab value %.0f ms %.3f e2e %.0f eval %.3f spmm frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["eval"]["ms"], d["roofline"]["frac"]))
print("fp32", d["fp32"]["ms_per_step"], "student", json.dumps(d.get("student"))[:900])
print("cpu", d.get("cpu_baseline"))
PY
tail -3 gpurun_out/bench_default.log | cut -c1-300
