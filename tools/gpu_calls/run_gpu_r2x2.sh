#!/bin/bash
# round-2 call X2 (8 GPUs): node-partitioned encoder, staged peer pull + pull of the scorer's rows + sparse gradient return
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29655"
timeout 500 $RUN tools/np_parity.py --peer --time > gpurun_out/np_peer_8gpu_v2.log 2>&1; echo "peer timing exit=$?"; grep "^{" gpurun_out/np_peer_8gpu_v2.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('ok', d['ok'], d['fp32']['ok'], d['bf16']['ok']); print(json.dumps(d.get('timing')))
"; grep -i "error\|Traceback" -A8 gpurun_out/np_peer_8gpu_v2.log | head -30
