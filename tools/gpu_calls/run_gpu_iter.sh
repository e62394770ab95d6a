#!/bin/bash
# iteration loop on the GPU box: full parity suite, then micro-benchmarks, then the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -n ${TAILN:-12} gpurun_out/pytest_gpu.log
timeout 600 python tools/kbench.py ${KB:-} > gpurun_out/kbench.log 2>&1; echo "kbench exit=$?"; cat gpurun_out/kbench.log | tail -40
timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.log 2>&1; echo "bench exit=$?"; tail -2 gpurun_out/bench.log
