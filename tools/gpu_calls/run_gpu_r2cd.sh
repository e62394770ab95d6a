#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/kbench.py gemmexp 2>&1 | grep -A1 "K=256+256" | cut -c1-260
