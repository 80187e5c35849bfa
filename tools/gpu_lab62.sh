#!/bin/bash
# ncu --set full of the shipped packed K4 (now compiled for 9 blocks per SM, 56 registers), 1M agents
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:step_boids_packed -s 40 -c 1 -o gpurun_out/lab62_k4_packed python tools/k4_ab.py --agents 1000000 --variants 0 --steps 5 --settle 30 > gpurun_out/lab62_ncu.log 2>&1
tail -2 gpurun_out/lab62_ncu.log
