#!/bin/bash
# round-2 call N: where does the student-step capture get invalidated inside pytest (line tracer), then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x -p tools.capture_trace -s --timeout=600 --timeout-method=thread > gpurun_out/t_trace.log 2>&1; echo "trace exit=$?"
grep -n "CAPTURE INVALIDATED" -A40 gpurun_out/t_trace.log | cut -c1-240 | head -120
tail -n 5 gpurun_out/t_trace.log | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --timeout-method=thread > gpurun_out/t_all.log 2>&1; echo "all exit=$?"; tail -n 12 gpurun_out/t_all.log | cut -c1-200
timeout 600 python tools/kbench.py spmmab > gpurun_out/kbench_spmmab.log 2>&1; cat gpurun_out/kbench_spmmab.log | cut -c1-200
