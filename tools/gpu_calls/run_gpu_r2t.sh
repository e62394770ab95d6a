#!/bin/bash
# round-2 call T: whole GPU suite + launch list of the bf16 and fp32 C4 steps with the streaming SpMM
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --timeout-method=thread > gpurun_out/t_all.log 2>&1; echo "all exit=$?"; tail -n 4 gpurun_out/t_all.log | cut -c1-200
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bf16.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list bf16 exit=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --precision fp32"
timeout 300 $CMD > gpurun_out/plain32.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_fp32.csv $CMD > gpurun_out/ncu_launches32.log 2>&1
echo "launch list fp32 exit=$?"
python tools/launch_breakdown.py gpurun_out/launches_bf16.csv 2>&1 | tail -40
