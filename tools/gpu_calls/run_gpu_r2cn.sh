#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -m gpu -q -x --timeout=600 > gpurun_out/t_k.log 2>&1; echo "kernel+e2e tests exit=$?"; tail -n 2 gpurun_out/t_k.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fp32 --no-student --no-hoisted > gpurun_out/bench_default.log 2>&1; echo "bench exit=$?"
grep "^{" gpurun_out/bench_default.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms %.4f e2e %.0f eval %.3f' % (d['ms_per_step'], d['e2e']['value'], d['eval']['ms']))"
