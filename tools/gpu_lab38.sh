#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_custom_step.py -x -q -m gpu > gpurun_out/lab38_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/lab38_pytest.log
tail -40 gpurun_out/lab38_pytest.log
