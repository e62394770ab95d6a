#!/bin/bash
# round-2 call W (N GPUs): peer-memory SpMM pieces in graph replay + strong-scaling step, peer vs all-gather
mkdir -p gpurun_out
NG=${NG:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NG --master-addr 127.0.0.1 --master-port 29655"
timeout 500 $RUN tools/np_parity.py --peer --time --comm > gpurun_out/np_peer_${NG}gpu.log 2>&1; echo "peer timing exit=$?"; grep "^{" gpurun_out/np_peer_${NG}gpu.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('ok', d['ok']); print(json.dumps(d.get('timing'))); [print(k, v) for k,v in d.get('pieces',{}).items()]
"; grep -i "error\|Traceback" -A8 gpurun_out/np_peer_${NG}gpu.log | head -30
if [ "$NG" != "2" ]; then
timeout 500 $RUN tools/np_parity.py --time --comm > gpurun_out/np_ag_${NG}gpu.log 2>&1; echo "all-gather timing exit=$?"; grep "^{" gpurun_out/np_ag_${NG}gpu.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('ok', d['ok']); print(json.dumps(d.get('timing'))); [print(k, v) for k,v in d.get('pieces',{}).items()]
"
fi
