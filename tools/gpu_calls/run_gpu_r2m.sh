#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=0 timeout 300 python tools/debug_capture.py fp32 > gpurun_out/debug_capture_fp32.log 2>&1; echo "debug fp32 exit=$?"; grep -n "INVALIDATED\|done\|Error\|error" gpurun_out/debug_capture_fp32.log | head -20; tail -25 gpurun_out/debug_capture_fp32.log | cut -c1-220
for st in 6 12; do for ct in 148 296; do timeout 120 tools/_build/gather4_bench $st $ct 1 2>&1 | grep -v "^ldg 4-in-flight, [0-9]* warps (1" ; done; done > gpurun_out/gather4_bench.log 2>&1; cat gpurun_out/gather4_bench.log
