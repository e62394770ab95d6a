#!/bin/bash
mkdir -p gpurun_out
for wl in collab-student cora-student; do
timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.log 2>&1; echo "bench $wl exit=$?"
grep "^{" gpurun_out/bench_$wl.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['config'].get('encoder_rows'))" || tail -5 gpurun_out/bench_$wl.log
done
