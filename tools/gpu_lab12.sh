#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_coltile.py -q -m gpu > gpurun_out/lab12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/lab12_pytest.log
KG_CT_STAGE=1 timeout 600 python -m pytest tests/test_gpu_coltile.py -q -m gpu >> gpurun_out/lab12_pytest.log 2>&1; echo "pytest(stage1) rc=$?" >> gpurun_out/lab12_pytest.log
tail -15 gpurun_out/lab12_pytest.log
{
timeout 300 python tools/k4_ab.py --agents 1000000 --variants 0,5
KG_CT_STAGE=1 timeout 300 python tools/k4_ab.py --agents 1000000 --variants 5
timeout 300 python tools/k4_ab.py --agents 1000000 --variants 0,5 --flush
timeout 300 python tools/k4_ab.py --agents 8000000 --variants 0,5 --steps 20
} > gpurun_out/lab12_ab.jsonl 2> gpurun_out/lab12_ab.err
cat gpurun_out/lab12_ab.jsonl; tail -5 gpurun_out/lab12_ab.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_boids_coltile -s 6 -c 1 -o gpurun_out/lab12_coltile python tools/k4_ab.py --agents 1000000 --variants 5 --steps 5 --settle 30 > gpurun_out/lab12_ncu.log 2>&1
tail -3 gpurun_out/lab12_ncu.log
