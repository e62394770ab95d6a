#!/bin/bash
# round-2 call CH: loop-invariant input aggregation kept on the graph: tests + default bench line (headline without it, extra object with it)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_config_sizes.py -m gpu -q -x --timeout=900 > gpurun_out/t_e2e.log 2>&1; echo "e2e+config tests exit=$?"; tail -n 3 gpurun_out/t_e2e.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fp32 --no-student > gpurun_out/bench_default.log 2>&1; echo "bench default exit=$?"
grep "^{" gpurun_out/bench_default.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); h=d['invariant_hoisted']
print('headline ms %.4f e2e %.0f eval %.3f spmm launches %d frac %.3f' % (d['ms_per_step'], d['e2e']['value'], d['eval']['ms'], d['roofline']['launches_timed'], d['roofline']['frac']))
print('hoisted  ms %.4f value %.0f e2e %.0f eval %.3f spmm launches %d frac %.3f' % (h['ms_per_step'], h['value'], h['e2e']['value'], h['eval']['ms'], h['roofline']['launches_timed'], h['roofline']['frac']))"
