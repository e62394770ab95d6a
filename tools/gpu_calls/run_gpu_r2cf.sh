#!/bin/bash
# round-2 call CF: feature-minibatch student step encodes every node once (when it touches >= N rows, dropout 0)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_config_sizes.py -m gpu -q -x --timeout=900 > gpurun_out/t_e2e.log 2>&1; echo "e2e+config tests exit=$?"; tail -n 3 gpurun_out/t_e2e.log | cut -c1-300
timeout 500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fp32 > gpurun_out/bench_default.log 2>&1; echo "bench default exit=$?"; grep "^{" gpurun_out/bench_default.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d[\"ms_per_step\"], d[\"student\"][\"ms_per_step\"], d[\"student\"][\"value\"])"
