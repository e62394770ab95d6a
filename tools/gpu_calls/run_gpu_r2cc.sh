#!/bin/bash
mkdir -p gpurun_out
for wl in physics-student cora-student; do timeout 300 python tools/student_host_profile.py $wl 2>&1 | tail -9; done
