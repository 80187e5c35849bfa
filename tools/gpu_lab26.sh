#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"strip_step_kernel|step_boids_packed" -s 16 -c 12 -o gpurun_out/lab26_strip python tools/strip_ncu_probe.py 2000000 > gpurun_out/lab26_ncu.log 2>&1
tail -3 gpurun_out/lab26_ncu.log
