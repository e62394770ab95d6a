#!/bin/bash
# round-2 call K: TF32 with six splitter warps + lean split: tests, accuracy, time; driver end-to-end tests
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tf32 or wgrad or gemm" --timeout=200 --timeout-method=thread > gpurun_out/t_k.log 2>&1; echo "kernels exit=$?"; tail -n 4 gpurun_out/t_k.log
timeout 300 python tools/kbench.py tf32 2>&1 | grep tf32x3 | cut -c1-150
timeout 300 python tools/fp32_accuracy.py c4 > gpurun_out/fp32_accuracy_c4.txt 2>&1; grep -v Warn gpurun_out/fp32_accuracy_c4.txt | cut -c1-72 | tail -17
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_config_sizes.py -m gpu -q --timeout=600 --timeout-method=thread > gpurun_out/t_e2e.log 2>&1; echo "e2e+cfg exit=$?"; tail -n 6 gpurun_out/t_e2e.log
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_collab.log 2>&1; echo "bench collab exit=$?"
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_collab.log") if x.startswith("{")][-1])
print("collab value %.0f ms %.3f e2e %.0f eval %.3f spmm frac %.3f share %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["eval"]["ms"], d["roofline"]["frac"], d["roofline"]["share_of_step"]))
f=d["fp32"]; print("fp32: value %.0f ms %.3f ratio %.2f dense %s" % (f["value"], f["ms_per_step"], f["ratio_to_bf16_step"], f["roofline"].get("dense_layers")))
PY
