#!/bin/bash
# round-2 call D: fp32 accuracy A/B (TF32x3 vs SIMT vs oracle vs fp64), integer-rounding splitter timing, new bench.py workloads
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tf32x3 or tf32 or wgrad or colsum" --timeout=120 --timeout-method=thread > gpurun_out/t_tf32.log 2>&1; echo "tf32 exit=$?"; tail -n 5 gpurun_out/t_tf32.log
timeout 600 python tools/fp32_accuracy.py > gpurun_out/fp32_accuracy.txt 2>&1; echo "acc exit=$?"; cat gpurun_out/fp32_accuracy.txt | tail -40
timeout 300 python tools/kbench.py tf32 > gpurun_out/kbench_tf32.log 2>&1; echo "kbench exit=$?"; grep tf32x3 gpurun_out/kbench_tf32.log
for wl in cora-student physics-student collab-student; do
  timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 --cpu-baseline-seconds 10 > gpurun_out/bench_$wl.log 2>&1; echo "bench $wl exit=$?"; tail -c 2200 gpurun_out/bench_$wl.log
done
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_collab.log 2>&1; echo "bench collab exit=$?"; tail -c 4500 gpurun_out/bench_collab.log
