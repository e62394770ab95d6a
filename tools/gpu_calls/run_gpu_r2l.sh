#!/bin/bash
# round-2 call L: capture-invalidation hunt in the student step (fp32), TF32 trunc-hi experiment
mkdir -p gpurun_out
timeout 300 python tools/debug_capture.py fp32 > gpurun_out/debug_capture_fp32.log 2>&1; echo "debug fp32 exit=$?"; grep -v Warning gpurun_out/debug_capture_fp32.log | tail -12 | cut -c1-300
timeout 300 python tools/debug_capture.py bf16 > gpurun_out/debug_capture_bf16.log 2>&1; echo "debug bf16 exit=$?"; tail -4 gpurun_out/debug_capture_bf16.log | cut -c1-300
LLP_TUNING=25=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tf32 or wgrad" --timeout=200 --timeout-method=thread > gpurun_out/t_trunc.log 2>&1; echo "trunc-hi tests exit=$?"; tail -n 12 gpurun_out/t_trunc.log | cut -c1-200
LLP_TUNING=25=1 timeout 300 python tools/kbench.py tf32 2>&1 | grep tf32x3 | cut -c1-150
LLP_TUNING=25=1 timeout 300 python tools/fp32_accuracy.py c4 2>&1 | grep -v Warn | cut -c1-72 | tail -17
