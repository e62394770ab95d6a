#!/bin/bash
# round-2 call G: full GPU suite (chunk=1 TF32 default, hub combine with selective fence, graph-captured eval), benches, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --timeout-method=thread > gpurun_out/t_all.log 2>&1; echo "all exit=$?"; tail -n 8 gpurun_out/t_all.log
timeout 300 python tools/kbench.py spmm > gpurun_out/kbench_spmm_merge.log 2>&1; cat gpurun_out/kbench_spmm_merge.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_collab.log 2>&1; echo "bench collab exit=$?"
for wl in cora-student physics-student collab-student coauthor-physics cora; do
  timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 --cpu-baseline-seconds 8 > gpurun_out/bench_$wl.log 2>&1; echo "bench $wl exit=$?"
done
python - <<'PY'
import json
for wl in ["collab","cora-student","physics-student","collab-student","coauthor-physics","cora"]:
    try:
        d=json.loads([x for x in open(f"gpurun_out/bench_{wl}.log") if x.startswith("{")][-1])
        print(wl, "value %.0f ms %.3f host %.3f e2e %.0f launches %d cpu %s" % (d["value"], d["ms_per_step"], d["host_enqueue_ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d.get("cpu_baseline",{}).get("value")))
        print("   roofline", {k:d["roofline"].get(k) for k in ("bound","achieved","peak","frac","share_of_step")}, "eval", d.get("eval",{}).get("ms"))
        if "fp32" in d:
            f=d["fp32"]; print("   fp32: value %.0f ms %.3f ratio %.2f dense %s" % (f["value"], f["ms_per_step"], f["ratio_to_bf16_step"], f["roofline"].get("dense_layers")))
    except Exception as e: print(wl, "ERR", repr(e))
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bf16.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"
python tools/launch_breakdown.py gpurun_out/launches_bf16.csv 2>&1 | tail -45
