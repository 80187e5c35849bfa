#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python tools/k4_ab.py --agents 20000 --variants 6 --steps 2 --settle 2 2>&1 | grep -v "^$" | head -40 > gpurun_out/lab14_memcheck.txt
head -20 gpurun_out/lab14_memcheck.txt
bash tools/gpu_lab13.sh
