#!/bin/bash
# ncu passes (launch list + one full capture of the top kernel), each only after the plain command exited 0.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-spmm_kernel} -s ${KSKIP:-10} -c ${KCOUNT:-5} -o gpurun_out/prof_${KTAG:-spmm} -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit=$?"
tail -3 gpurun_out/plain.log
