#!/bin/bash
# round-2 call AC: fused edge scorer with 16 gather warps (default build) — parity tests, micro-benchmark, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -m gpu -q -k "edge or scorer or fused or teacher or captured" --timeout=400 --timeout-method=thread 2>&1 | tail -3
timeout 300 python tools/kbench.py edgemlp 2>&1 | grep "fused edge" | cut -c1-160
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fp32 --no-student > gpurun_out/bench_em16.log 2>&1
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_em16.log") if x.startswith("{")][-1])
print("16 gather warps: collab value %.0f ms %.3f eval %.3f" % (d["value"], d["ms_per_step"], d["eval"]["ms"]))
PY
