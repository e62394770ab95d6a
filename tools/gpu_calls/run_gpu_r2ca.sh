#!/bin/bash
# round-2 call CA (2 GPUs): multi-GPU pytest + the N = 2 bench line with the final kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=600 > gpurun_out/t_multi.log 2>&1; echo "multi exit=$?"; tail -n 3 gpurun_out/t_multi.log | cut -c1-200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 exit=$?"
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_n2.log") if x.startswith("{")][-1])
print("N=2 value %.0f ms %.4f e2e %.0f eval %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["eval"]["ms"]))
PY
