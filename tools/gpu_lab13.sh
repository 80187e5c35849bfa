#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_field2d.py tests/test_golden.py -q -m gpu -k "variant or numpy" > gpurun_out/lab13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/lab13_pytest.log
tail -15 gpurun_out/lab13_pytest.log
{
timeout 300 python tools/k4_ab.py --agents 1000000 --variants 0,6,0,6 --check
timeout 300 python tools/k4_ab.py --agents 1000000 --variants 0,6 --flush
timeout 300 python tools/k4_ab.py --agents 8000000 --variants 0,6 --steps 20
} > gpurun_out/lab13_ab.jsonl 2> gpurun_out/lab13_ab.err
cat gpurun_out/lab13_ab.jsonl; tail -5 gpurun_out/lab13_ab.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_boids_staged -s 6 -c 1 -o gpurun_out/lab13_staged python tools/k4_ab.py --agents 1000000 --variants 6 --steps 5 --settle 30 > gpurun_out/lab13_ncu.log 2>&1
tail -3 gpurun_out/lab13_ncu.log
