#!/bin/bash
# GPU lab call 10 (round 2, session 2): the column-chunk K4 (KG_K4_COLTILE = 5) — parity, A/B, one ncu capture.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_coltile.py -x -q -m gpu > gpurun_out/lab10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/lab10_pytest.log
tail -15 gpurun_out/lab10_pytest.log
{
for fl in "" "--flush"; do
  timeout 300 python tools/k4_ab.py --agents 1000000 --variants 0,5,4 $fl
done
timeout 300 python tools/k4_ab.py --agents 8000000 --variants 0,5 --steps 20
} > gpurun_out/lab10_ab.jsonl 2> gpurun_out/lab10_ab.err
cat gpurun_out/lab10_ab.jsonl; tail -5 gpurun_out/lab10_ab.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_boids_coltile -s 40 -c 2 -o gpurun_out/lab10_coltile python tools/k4_ab.py --agents 1000000 --variants 5 --steps 5 --settle 30 > gpurun_out/lab10_ncu.log 2>&1
tail -3 gpurun_out/lab10_ncu.log
