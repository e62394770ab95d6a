#!/bin/bash
# round-2 call A: TF32x3 kernel tests first (own timeout), new config-size parity tests, whole GPU suite, bf16 + fp32 bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tf32x3 or tf32" --timeout=120 --timeout-method=thread > gpurun_out/t_tf32.log 2>&1; echo "tf32 exit=$?"; tail -n 30 gpurun_out/t_tf32.log
timeout 900 python -m pytest tests/test_gpu_config_sizes.py -m gpu -q --timeout=600 --timeout-method=thread > gpurun_out/t_cfg.log 2>&1; echo "cfg exit=$?"; tail -n 40 gpurun_out/t_cfg.log
timeout 600 python -m pytest tests -m gpu -q --timeout=300 --timeout-method=thread --deselect tests/test_gpu_config_sizes.py > gpurun_out/t_all.log 2>&1; echo "all exit=$?"; tail -n 12 gpurun_out/t_all.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_bf16.log 2>&1; echo "bench bf16 exit=$?"; tail -c 3000 gpurun_out/bench_bf16.log
timeout 300 python bench.py --steps 10 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/bench_fp32.log 2>&1; echo "bench fp32 exit=$?"; tail -c 2500 gpurun_out/bench_fp32.log
