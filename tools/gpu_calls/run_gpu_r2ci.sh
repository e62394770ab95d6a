#!/bin/bash
# round-2 call CI: sanity of the final kernels on the 10M-node power-law workload (configs[4]) at full scale, one GPU
mkdir -p gpurun_out
timeout 800 python bench.py --workload powerlaw-10m --scale 1.0 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_powerlaw.log 2>&1; echo "bench powerlaw exit=$?"
grep "^{" gpurun_out/bench_powerlaw.log | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print('nodes %d messages %d ms %.3f value %.0f e2e %.0f frac %.3f share %.3f eval %s' % (d['config']['nodes'], d['config']['messages'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['share_of_step'], (d.get('eval') or {}).get('ms')))" || tail -5 gpurun_out/bench_powerlaw.log
