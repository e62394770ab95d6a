#!/bin/bash
# round-2 call C: config-size tests (fp64-truth gradient bounds), TF32 micro-benchmarks, fp32 step launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_config_sizes.py -m gpu -q --timeout=600 --timeout-method=thread > gpurun_out/t_cfg.log 2>&1; echo "cfg exit=$?"; tail -n 30 gpurun_out/t_cfg.log
timeout 300 python tools/kbench.py tf32 > gpurun_out/kbench_tf32.log 2>&1; echo "kbench exit=$?"; cat gpurun_out/kbench_tf32.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --precision fp32"
timeout 300 $CMD > gpurun_out/plain_fp32.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_fp32.csv $CMD > gpurun_out/ncu_launches_fp32.log 2>&1
echo "launch list exit=$?"
python tools/launch_breakdown.py gpurun_out/launches_fp32.csv 2>&1 | tail -40
