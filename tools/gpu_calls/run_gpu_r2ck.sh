#!/bin/bash
# round-2 call CK: whole GPU suite + smoke after the pinning changes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --timeout-method=thread > gpurun_out/t_all.log 2>&1; echo "all exit=$?"; tail -n 3 gpurun_out/t_all.log | cut -c1-200
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -n 1 gpurun_out/smoke.log | cut -c1-200
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fp32 --no-student > gpurun_out/bench_default.log 2>&1; echo "bench exit=$?"
grep "^{" gpurun_out/bench_default.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); h=d['invariant_hoisted']
print('headline ms %.4f e2e %.0f eval %.3f %s' % (d['ms_per_step'], d['e2e']['value'], d['eval']['ms'], [round(x,3) for x in d['eval']['ms_of_5_passes']]))
print('hoisted  ms %.4f value %.0f e2e %.0f eval %.3f' % (h['ms_per_step'], h['value'], h['e2e']['value'], h['eval']['ms']))"
