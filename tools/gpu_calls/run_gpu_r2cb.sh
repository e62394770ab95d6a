#!/bin/bash
# round-2 call CB: speculative candidate draw for the dense negative sampling: parity tests + the host-bound workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=600 -k "negative" > gpurun_out/t_neg.log 2>&1; echo "neg tests exit=$?"; tail -n 2 gpurun_out/t_neg.log | cut -c1-300
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_config_sizes.py -m gpu -q -x --timeout=900 > gpurun_out/t_e2e.log 2>&1; echo "e2e+config tests exit=$?"; tail -n 2 gpurun_out/t_e2e.log | cut -c1-300
for wl in cora-student physics-student coauthor-physics cora; do
  timeout 400 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.log 2>&1; echo "bench $wl exit=$?"
done
python - <<'PY'
import json
for wl in ["cora-student","physics-student","coauthor-physics","cora"]:
    try:
        d=json.loads([x for x in open(f"gpurun_out/bench_{wl}.log") if x.startswith("{")][-1])
        print(wl, "value %.0f ms %.3f host %s e2e %.0f (%.3f ms) launches %s" % (d["value"], d["ms_per_step"], d.get("host_enqueue_ms_per_step"), d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("gpu_launches")))
    except Exception as e: print(wl, "ERR", repr(e))
PY
