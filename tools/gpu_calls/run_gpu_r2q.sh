#!/bin/bash
# round-2 call Q: streaming SpMM v3 (fast path + fall-through pieces): parity tests, A/B timing, instruction count, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_config_sizes.py -m gpu -q -k "spmm or c4 or c5 or sage" --timeout=400 --timeout-method=thread > gpurun_out/t_spmm.log 2>&1; echo "spmm tests exit=$?"; tail -n 5 gpurun_out/t_spmm.log | cut -c1-200
timeout 600 python tools/kbench.py spmmab > gpurun_out/kbench_spmmab3.log 2>&1; cat gpurun_out/kbench_spmmab3.log | cut -c1-200
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fp32 > gpurun_out/bench_collab_q.log 2>&1; echo "bench exit=$?"
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_collab_q.log") if x.startswith("{")][-1])
print("collab value %.0f ms %.3f e2e %.0f eval %.3f spmm frac %.3f share %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["eval"]["ms"], d["roofline"]["frac"], d["roofline"]["share_of_step"]))
PY
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:spmm_stream -c 6 --csv --log-file gpurun_out/ncu_stream_inst.csv python tools/spmm_only.py > /dev/null 2>&1; grep -v "^==" gpurun_out/ncu_stream_inst.csv | cut -d, -f5,13- | cut -c1-200
