#!/bin/bash
# round-2 call P: ncu --set full of the streaming and the row-run SpMM (bf16 F=256, cold L2), stall reasons
mkdir -p gpurun_out
timeout 200 python tools/spmm_only.py > gpurun_out/plain_spmm.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_stream_kernel -c 2 -o gpurun_out/prof_spmm_stream_r02 -f python tools/spmm_only.py > gpurun_out/ncu_spmm_stream.log 2>&1
echo "stream capture exit=$?"
LLP_TUNING=3=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -c 2 -o gpurun_out/prof_spmm_rowrun_r02 -f python tools/spmm_only.py > gpurun_out/ncu_spmm_rowrun.log 2>&1
echo "row-run capture exit=$?"
ls -la gpurun_out/*.ncu-rep
