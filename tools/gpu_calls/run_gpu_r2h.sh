#!/bin/bash
# round-2 call H: TF32 k-block 16 (SWIZZLE_64B, 4 stages) vs 32 (2 stages): correctness, accuracy, time; SpMM after the merge revert
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=200 --timeout-method=thread > gpurun_out/t_k.log 2>&1; echo "kernels exit=$?"; tail -n 6 gpurun_out/t_k.log
LLP_TUNING=24=32 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tf32 or wgrad" --timeout=200 --timeout-method=thread > gpurun_out/t_k32.log 2>&1; echo "kernels bk32 exit=$?"; tail -n 3 gpurun_out/t_k32.log
for bk in 16 32; do
  echo "=== k-block $bk"
  LLP_TUNING=24=$bk timeout 300 python tools/kbench.py tf32 2>&1 | grep tf32x3 > gpurun_out/kbench_tf32_bk$bk.log; cut -c1-150 gpurun_out/kbench_tf32_bk$bk.log
done
timeout 300 python tools/fp32_accuracy.py > gpurun_out/fp32_accuracy.txt 2>&1; grep -v Warn gpurun_out/fp32_accuracy.txt | cut -c1-72
timeout 300 python tools/kbench.py spmm > gpurun_out/kbench_spmm.log 2>&1; cat gpurun_out/kbench_spmm.log
timeout 900 python -m pytest tests/test_gpu_config_sizes.py tests/test_gpu_e2e.py -m gpu -q --timeout=600 --timeout-method=thread > gpurun_out/t_cfg.log 2>&1; echo "cfg+e2e exit=$?"; tail -n 6 gpurun_out/t_cfg.log
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_collab.log 2>&1; echo "bench collab exit=$?"
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_collab.log") if x.startswith("{")][-1])
print("collab value %.0f ms %.3f e2e %.0f eval %.3f spmm frac %.3f share %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["eval"]["ms"], d["roofline"]["frac"], d["roofline"]["share_of_step"]))
f=d["fp32"]; print("fp32: value %.0f ms %.3f ratio %.2f dense %s" % (f["value"], f["ms_per_step"], f["ratio_to_bf16_step"], f["roofline"].get("dense_layers")))
PY
