#!/bin/bash
# round-2 call AJ (2 GPUs): multi-GPU pytest after the teardown change (peer mappings closed in finish_distributed)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=800 --timeout-method=thread > gpurun_out/t_multi.log 2>&1; echo "multi exit=$?"; tail -n 4 gpurun_out/t_multi.log | cut -c1-300
