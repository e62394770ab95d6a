#!/bin/bash
# round-2 call BH: tall-tile NT GEMM (256 rows per weight pass) vs the streaming kernel: bit equality, timing, step A/B
mkdir -p gpurun_out
timeout 300 python tools/kbench.py tall 2>&1 | grep -v "^$" | cut -c1-220
for rep in 1 2; do
for knob in "" "20=4"; do
  LLP_TUNING=$knob timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-fp32 --no-student --allow-tuning > gpurun_out/bench_tall.log 2>&1
  python - "$knob" <<'PY'
import json,sys
d=json.loads([x for x in open("gpurun_out/bench_tall.log") if x.startswith("{")][-1])
print("knob '%s' ms %.4f e2e %.4f ms eval %.3f" % (sys.argv[1], d["ms_per_step"], d["e2e"]["ms_per_step"], d["eval"]["ms"]))
PY
done
done
