#!/bin/bash
# One gpurun call: staged GPU checks, each under its own timeout so a hung kernel cannot eat the whole budget.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
stage() { name=$1; shift; echo "=== $name"; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?"; tail -n ${TAILN:-15} gpurun_out/$name.log; }
stage k_simt python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not tcgen05" --timeout=200 --timeout-method=thread -x
stage k_tc python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tcgen05" --timeout=200 --timeout-method=thread
stage e2e_fp32 python -m pytest tests/test_gpu_e2e.py -m gpu -q -k "fp32 or minibatch" --timeout=200 --timeout-method=thread
stage e2e_bf16 python -m pytest tests/test_gpu_e2e.py -m gpu -q -k "not fp32 and not minibatch" --timeout=200 --timeout-method=thread
stage smoke python -c "import __graft_entry__ as g; g.smoke()"
TAILN=3 stage bench python bench.py --steps 5 --warmup 3
