#!/bin/bash
# round-2 call R: software-pipelined SpMM (pipe) vs streaming vs row-run
mkdir -p gpurun_out
timeout 600 python tools/kbench.py spmmab > gpurun_out/kbench_spmmab4.log 2>&1; grep -v "variant [13]" gpurun_out/kbench_spmmab4.log | cut -c1-200
for m in 3 4; do
LLP_TUNING=3=$m timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fp32 --allow-tuning > gpurun_out/bench_collab_r$m.log 2>&1; echo "bench exit=$?"
python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/bench_collab_r$m.log") if x.startswith("{")][-1])
print("mode $m: collab value %.0f ms %.3f e2e %.0f eval %.3f spmm frac %.3f share %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["eval"]["ms"], d["roofline"]["frac"], d["roofline"]["share_of_step"]))
PY
done
LLP_TUNING=3=3 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_config_sizes.py -m gpu -q -k "spmm or c4 or c5 or sage" --timeout=400 --timeout-method=thread > gpurun_out/t_spmm_pipe.log 2>&1; echo "spmm tests (pipe) exit=$?"; tail -n 5 gpurun_out/t_spmm_pipe.log | cut -c1-200
