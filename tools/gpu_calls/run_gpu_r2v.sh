#!/bin/bash
# round-2 call V (2 GPUs): node-partitioned encoder over peer memory: parity, pieces, strong-scaling step
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29655"
timeout 300 $RUN tools/np_parity.py --peer > gpurun_out/np_peer_parity_2gpu.log 2>&1; echo "peer parity exit=$?"; grep "^{" gpurun_out/np_peer_parity_2gpu.log | cut -c1-1500; grep -i "error\|Traceback" -A8 gpurun_out/np_peer_parity_2gpu.log | head -40
timeout 400 $RUN tools/np_parity.py --peer --time --comm > gpurun_out/np_peer_2gpu.log 2>&1; echo "peer timing exit=$?"; grep "^{" gpurun_out/np_peer_2gpu.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(json.dumps(d.get('timing'),indent=0)); print(json.dumps(d.get('pieces'),indent=0))
"; grep -i "error\|Traceback" -A8 gpurun_out/np_peer_2gpu.log | head -30
timeout 400 $RUN tools/np_parity.py --time --comm > gpurun_out/np_ag_2gpu.log 2>&1; echo "all-gather timing exit=$?"; grep "^{" gpurun_out/np_ag_2gpu.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(json.dumps(d.get('timing'),indent=0)); print(json.dumps(d.get('pieces'),indent=0))
"
