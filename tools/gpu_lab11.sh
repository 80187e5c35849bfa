#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_boids_coltile -s 6 -c 1 -o gpurun_out/lab11_coltile python tools/k4_ab.py --agents 1000000 --variants 5 --steps 5 --settle 30 > gpurun_out/lab11_ncu.log 2>&1
tail -3 gpurun_out/lab11_ncu.log
