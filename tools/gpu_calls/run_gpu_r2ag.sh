#!/bin/bash
# round-2 call AG: probe — two-pass half-warp-per-row gather vs one-pass full-warp gather (cold L2)
mkdir -p gpurun_out
timeout 120 tools/_build/pairgather_bench 0 > gpurun_out/pairgather.log 2>&1; timeout 120 tools/_build/pairgather_bench 1 >> gpurun_out/pairgather.log 2>&1; cat gpurun_out/pairgather.log
