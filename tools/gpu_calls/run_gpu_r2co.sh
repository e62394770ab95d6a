#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload cora --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain_cora.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_cora.csv $CMD > gpurun_out/ncu_cora.log 2>&1
echo "launch list cora exit=$?"
python tools/launch_breakdown.py gpurun_out/launches_cora.csv 2>&1 | head -30
python tools/launch_breakdown.py gpurun_out/launches_cora.csv 2>&1 | tail -1
