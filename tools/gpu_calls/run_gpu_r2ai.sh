#!/bin/bash
# round-2 call AI (4 GPUs): the driver's scaling launch at N = 4
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29693 bench.py --gpus 4 --steps 20 --warmup 3 --no-fp32 > gpurun_out/bench_n4.log 2>&1; echo "bench n4 exit=$?"
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_n4.log") if x.startswith("{")][-1])
print("N=4 collab value %.0f ms %.3f e2e %.0f (%.3f ms) eval %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["eval"]["ms"]))
PY
