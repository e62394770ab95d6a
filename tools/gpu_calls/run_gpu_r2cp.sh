#!/bin/bash
# round-2 call CP: SpMM over the padded width for odd feature widths (Cora's 1,433): tests + the cora teacher bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_config_sizes.py tests/test_gpu_e2e.py -m gpu -q -x --timeout=900 > gpurun_out/t_k.log 2>&1; echo "tests exit=$?"; tail -n 2 gpurun_out/t_k.log | cut -c1-200
timeout 300 python bench.py --workload cora --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cora.log 2>&1; echo "bench cora exit=$?"
grep "^{" gpurun_out/bench_cora.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cora ms %.4f value %.0f e2e %.0f eval %.3f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['eval']['ms']))"
