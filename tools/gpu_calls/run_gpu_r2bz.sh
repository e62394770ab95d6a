#!/bin/bash
# round-2 call BZ: final single-GPU evidence of this session's kernels: whole suite, all bench workloads, launch lists,
# ncu --set full of the SpMM and of the fused edge scorer
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --timeout-method=thread > gpurun_out/t_all.log 2>&1; echo "all exit=$?"; tail -n 3 gpurun_out/t_all.log | cut -c1-200
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -n 2 gpurun_out/smoke.log | cut -c1-200
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32 --no-student --no-hoisted"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bf16.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list bf16 exit=$?"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spmm|edge_mlp" -s 30 -c 12 -o gpurun_out/prof_spmm_edge_r02b -f $CMD > gpurun_out/ncu_spmm.log 2>&1
echo "full capture spmm + edge_mlp exit=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --precision fp32 --no-student --no-hoisted"
timeout 300 $CMD > gpurun_out/plain32.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_fp32.csv $CMD > gpurun_out/ncu_launches32.log 2>&1
echo "launch list fp32 exit=$?"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_collab.log 2>&1; echo "bench collab exit=$?"
for wl in cora-student physics-student collab-student coauthor-physics cora; do
  timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 --cpu-baseline-seconds 8 > gpurun_out/bench_$wl.log 2>&1; echo "bench $wl exit=$?"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "bench reference exit=$?"
python - <<'PY'
import json
for wl in ["collab","cora-student","physics-student","collab-student","coauthor-physics","cora","reference"]:
    try:
        d=json.loads([x for x in open(f"gpurun_out/bench_{wl}.log") if x.startswith("{")][-1])
        print(wl, "value %.0f ms %.3f host %s e2e %.0f launches %s cpu %s" % (d["value"], d["ms_per_step"], d.get("host_enqueue_ms_per_step"), d["e2e"]["value"], d.get("gpu_launches"), d.get("cpu_baseline",{}).get("value")))
        if "roofline" in d: print("   roofline", {k:d["roofline"].get(k) for k in ("bound","achieved","peak","frac","share_of_step")}, "eval", (d.get("eval") or {}).get("ms"))
        if "fp32" in d:
            f=d["fp32"]; print("   fp32: value %.0f ms %.3f ratio %.2f dense %s" % (f["value"], f["ms_per_step"], f["ratio_to_bf16_step"], f["roofline"].get("dense_layers")))
        if "student" in d:
            f=d["student"]; print("   student: value %.0f ms %.3f e2e %.0f frac %.3f" % (f["value"], f["ms_per_step"], f["e2e"]["value"], f["roofline"]["frac"]))
        if "invariant_hoisted" in d:
            f=d["invariant_hoisted"]; print("   hoisted: value %.0f ms %.4f e2e %.0f eval %.3f" % (f["value"], f["ms_per_step"], f["e2e"]["value"], f["eval"]["ms"]))
    except Exception as e: print(wl, "ERR", repr(e))
PY
python tools/launch_breakdown.py gpurun_out/launches_bf16.csv 2>&1 | tail -14
