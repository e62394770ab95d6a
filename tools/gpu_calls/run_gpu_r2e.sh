#!/bin/bash
# round-2 call E: TF32 chunk-length experiment (accuracy + time), SpMM in-kernel hub combine, host sampling speed-ups, full suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --timeout-method=thread > gpurun_out/t_all.log 2>&1; echo "all exit=$?"; tail -n 14 gpurun_out/t_all.log
for ch in 1 2 4; do
  echo "=== chunk_kb=$ch"
  LLP_TUNING=23=$ch timeout 300 python tools/fp32_accuracy.py c4 > gpurun_out/fp32_accuracy_chunk$ch.txt 2>&1; grep -v Warning gpurun_out/fp32_accuracy_chunk$ch.txt | tail -17 | cut -c1-70
  LLP_TUNING=23=$ch timeout 300 python tools/kbench.py tf32 2>&1 | grep tf32x3 > gpurun_out/kbench_tf32_chunk$ch.log; cat gpurun_out/kbench_tf32_chunk$ch.log | cut -c1-150
done
timeout 300 python tools/kbench.py spmm > gpurun_out/kbench_spmm_merge.log 2>&1; cat gpurun_out/kbench_spmm_merge.log
LLP_TUNING=22=1 timeout 300 python tools/kbench.py spmm > gpurun_out/kbench_spmm_nomerge.log 2>&1; cat gpurun_out/kbench_spmm_nomerge.log
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_collab.log 2>&1; echo "bench collab exit=$?"; tail -c 1500 gpurun_out/bench_collab.log | cut -c1-1500
LLP_TUNING=22=1 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fp32 --allow-tuning > gpurun_out/bench_collab_nomerge.log 2>&1; echo "bench collab nomerge exit=$?"; head -c 400 gpurun_out/bench_collab_nomerge.log
for wl in cora-student physics-student collab-student; do
  timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.log 2>&1; echo "bench $wl exit=$?"; head -c 420 gpurun_out/bench_$wl.log; echo
done
