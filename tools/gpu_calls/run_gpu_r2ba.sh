#!/bin/bash
# round-2 call BA: own counting sort for the edge plan + two-groups-in-flight edge scorer producer
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=600 -k "edge or hadamard or scorer or plan" > gpurun_out/t_edge.log 2>&1; echo "edge tests exit=$?"; tail -n 4 gpurun_out/t_edge.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x --timeout=600 > gpurun_out/t_e2e.log 2>&1; echo "e2e tests exit=$?"; tail -n 4 gpurun_out/t_e2e.log | cut -c1-300
timeout 300 python tools/kbench.py edgemlp misc 2>&1 | grep -i "edge\|MMA issue\|plan" | cut -c1-250
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fp32 > gpurun_out/bench_ba.log 2>&1; echo "bench exit=$?"
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_ba.log") if x.startswith("{")][-1])
print("collab value %.0f ms %.4f e2e %.0f (%.4f ms) launches %s eval %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("gpu_launches"), d["eval"]["ms"]))
PY
