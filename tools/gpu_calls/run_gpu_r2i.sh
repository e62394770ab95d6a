#!/bin/bash
# round-2 call I: ncu --set full captures (SpMM in the bf16 step, 3xTF32 GEMM in the fp32 step, bf16 GEMM / edge / wgrad kernels),
# device-side generation of the power-law workload at small scale, remaining tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_config_sizes.py tests/test_gpu_e2e.py -m gpu -q --timeout=600 --timeout-method=thread -k "c4 or c5 or norm" > gpurun_out/t_cfg.log 2>&1; echo "cfg+norm exit=$?"; tail -n 5 gpurun_out/t_cfg.log
timeout 300 python bench.py --workload powerlaw-10m --scale 0.05 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pl005.log 2>&1; echo "powerlaw 0.05 exit=$?"; tail -c 600 gpurun_out/bench_pl005.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -s 10 -c 5 -o gpurun_out/prof_spmm_r02 -f $CMD > gpurun_out/ncu_spmm.log 2>&1
echo "ncu spmm exit=$?"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel|gemm_nt_resb|edge_mlp_kernel|wgrad_kernel" -s 22 -c 11 -o gpurun_out/prof_dense_r02 -f $CMD > gpurun_out/ncu_dense.log 2>&1
echo "ncu dense exit=$?"
CMD32="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --precision fp32"
timeout 300 $CMD32 > gpurun_out/plain32.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 28 -c 6 -o gpurun_out/prof_tf32_r02 -f $CMD32 > gpurun_out/ncu_tf32.log 2>&1
echo "ncu tf32 exit=$?"
ls -la gpurun_out/*.ncu-rep
