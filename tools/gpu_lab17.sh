#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_closures.py tests/test_gpu_grid.py -x -q -m gpu > gpurun_out/lab17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/lab17_pytest.log
tail -40 gpurun_out/lab17_pytest.log
