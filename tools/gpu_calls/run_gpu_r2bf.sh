#!/bin/bash
# round-2 call BF: launch list of the collab student step (where do the non-GEMM 36 % go?)
mkdir -p gpurun_out
CMD="python bench.py --workload collab-student --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain_student.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_student.csv $CMD > gpurun_out/ncu_student.log 2>&1
echo "launch list exit=$?"
python tools/launch_breakdown.py gpurun_out/launches_student.csv 2>&1 | tail -45
