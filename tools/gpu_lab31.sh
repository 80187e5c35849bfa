#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_differential.py -x -q -m gpu -k "block" > gpurun_out/lab31_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/lab31_pytest.log
tail -30 gpurun_out/lab31_pytest.log
KG_FUZZ=1 timeout 1500 python -m pytest tests/test_gpu_differential.py -x -q -m gpu > gpurun_out/lab31_fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/lab31_fuzz.log
tail -30 gpurun_out/lab31_fuzz.log
