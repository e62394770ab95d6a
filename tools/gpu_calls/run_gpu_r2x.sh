#!/bin/bash
# round-2 call X (8 GPUs): node-partitioned encoder, staged peer pull vs all-gather: pieces + strong-scaling step;
# then the data-parallel bench at N = 8
mkdir -p gpurun_out
NG=8 tools/gpu_calls/run_gpu_r2w.sh
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 8 --steps 20 --warmup 3 --no-fp32 --no-cpu-baseline > gpurun_out/bench_n8.log 2>&1; echo "bench n8 exit=$?"
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_n8.log") if x.startswith("{")][-1])
print("N=8 collab value %.0f ms %.3f e2e %.0f (%.3f ms) eval %.3f spmm frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["eval"]["ms"], d["roofline"]["frac"]))
PY
