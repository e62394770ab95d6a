#!/bin/bash
# round-2 call B: TF32x3 v2 (chunked promotion, MN-major 32B-atom layout), fused KD loss, student graph step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tf32x3 or tf32 or llp_fused" --timeout=120 --timeout-method=thread > gpurun_out/t_tf32.log 2>&1; echo "tf32 exit=$?"; tail -n 25 gpurun_out/t_tf32.log
timeout 900 python -m pytest tests/test_gpu_config_sizes.py -m gpu -q --timeout=600 --timeout-method=thread > gpurun_out/t_cfg.log 2>&1; echo "cfg exit=$?"; tail -n 40 gpurun_out/t_cfg.log
timeout 600 python -m pytest tests -m gpu -q --timeout=300 --timeout-method=thread --deselect tests/test_gpu_config_sizes.py > gpurun_out/t_all.log 2>&1; echo "all exit=$?"; tail -n 25 gpurun_out/t_all.log
timeout 300 python bench.py --steps 10 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/bench_fp32.log 2>&1; echo "bench fp32 exit=$?"; tail -c 1500 gpurun_out/bench_fp32.log
