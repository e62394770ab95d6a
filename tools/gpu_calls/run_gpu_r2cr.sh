#!/bin/bash
# round-2 call CR: warp-per-row ordering of mid-size rows in the edge plan: parity tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -m gpu -q -x --timeout=500 > gpurun_out/t_k.log 2>&1; echo "tests exit=$?"; tail -n 2 gpurun_out/t_k.log | cut -c1-200
