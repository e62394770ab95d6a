#!/bin/bash
# round-2 call CM (8 GPUs): the default bench line at N = 8 with the final code
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/bench_n$NG.log 2>&1; echo "bench n$NG exit=$?"
grep "^{" gpurun_out/bench_n$NG.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=%d value %.0f ms %.4f e2e %.0f eval %.3f clocks %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['eval']['ms'], d['clocks']))" || tail -20 gpurun_out/bench_n$NG.log
