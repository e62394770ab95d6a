#!/bin/bash
# round-2 call J (8 GPUs): collab-shaped teacher at N = 8 (weak scaling, with the fp32 object), BASELINE.json configs[4] as written
# (10M nodes / 200M messages, edge-batch-sharded training + all-pairs-negative Hits@K at 8 x B200), N1 timing + link utilisation
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt; free -g | head -2 >> gpurun_out/gpus8.txt; nproc >> gpurun_out/gpus8.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.log 2>&1; echo "collab n8 exit=$?"
grep '^{' gpurun_out/bench_n8.log | cut -c1-700
timeout 1200 $TR --master-port 29532 bench.py --gpus 8 --workload powerlaw-10m --steps 10 --warmup 3 --no-fp32 > gpurun_out/bench_c5_n8.log 2>&1; echo "c5 n8 exit=$?"
grep '^{' gpurun_out/bench_c5_n8.log | cut -c1-900; tail -3 gpurun_out/bench_c5_n8.log | cut -c1-300
timeout 500 $TR --master-port 29533 tools/np_parity.py --time --comm > gpurun_out/np_parity_8gpu.log 2>&1; echo "np parity exit=$?"
grep '^{' gpurun_out/np_parity_8gpu.log | cut -c1-1500
