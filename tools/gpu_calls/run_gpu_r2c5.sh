#!/bin/bash
# round-2 call C5 (8 GPUs): BASELINE.json configs[4] — 10M nodes / 200M messages, edge-batch-sharded training + sharded Hits@K
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29681 bench.py --gpus 8 --workload powerlaw-10m --steps 5 --warmup 3 --no-fp32 --no-cpu-baseline > gpurun_out/bench_c5_n8.log 2>&1; echo "bench c5 n8 exit=$?"
grep "^{" gpurun_out/bench_c5_n8.log | tail -1 | cut -c1-3000; grep -i "error\|Traceback" -A6 gpurun_out/bench_c5_n8.log | head -30
free -g | head -2; df -h /dev/shm | tail -1
