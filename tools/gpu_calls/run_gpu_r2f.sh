#!/bin/bash
# round-2 call F (2 GPUs): multi-GPU parity inside pytest, real NCCL teardown, clock-sampler A/B for the DP step
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=800 --timeout-method=thread > gpurun_out/t_multi.log 2>&1; echo "multi exit=$?"; tail -n 15 gpurun_out/t_multi.log
for mode in none thread proc; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 --no-fp32 --clock-sampler $mode > gpurun_out/bench_n2_$mode.log 2>&1
  echo "bench n2 $mode exit=$?"; grep '^{' gpurun_out/bench_n2_$mode.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  value %.0f  ms/step %.4f  e2e ms %.4f  eval ms %.3f  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['eval']['ms'], d['clocks']))"
done
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 3 --no-fp32 --no-cpu-baseline --clock-sampler none > gpurun_out/bench_n1_none.log 2>&1; grep '^{' gpurun_out/bench_n1_none.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('N=1 none: value %.0f  ms/step %.4f  e2e ms %.4f eval ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['eval']['ms']))"
tail -5 gpurun_out/bench_n2_none.log | cut -c1-300
