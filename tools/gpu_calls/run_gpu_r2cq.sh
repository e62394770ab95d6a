#!/bin/bash
mkdir -p gpurun_out
for wl in coauthor-physics physics-student; do
CMD="python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$wl.csv $CMD > gpurun_out/ncu_$wl.log 2>&1
echo "launch list $wl exit=$?"
python tools/launch_breakdown.py gpurun_out/launches_$wl.csv 2>&1 | head -12
python tools/launch_breakdown.py gpurun_out/launches_$wl.csv 2>&1 | tail -1
done
