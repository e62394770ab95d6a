#!/bin/bash
# round-2 call AF (2 GPUs): the driver's scaling launch at N = 2 (both arms) on the final tree
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29691 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 exit=$?"
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_n2.log") if x.startswith("{")][-1])
print("N=2 collab value %.0f ms %.3f e2e %.0f (%.3f ms) eval %.3f keys %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["eval"]["ms"], sorted(d.keys())))
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29692 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | grep "^{" | cut -c1-300; echo "reference n2 exit=$?"
