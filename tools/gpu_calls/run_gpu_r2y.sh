#!/bin/bash
# round-2 call Y (2 GPUs): multi-GPU pytest (DP parity, node-partitioned all-gather / staged pull / per-edge loads) + W=2 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=800 --timeout-method=thread > gpurun_out/t_multi.log 2>&1; echo "multi exit=$?"; tail -n 15 gpurun_out/t_multi.log | cut -c1-300
NG=2 tools/gpu_calls/run_gpu_r2w.sh 2>&1 | head -8 | cut -c1-400
