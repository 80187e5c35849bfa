#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blocks.py -x -q -m gpu > gpurun_out/lab21_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/lab21_pytest.log
tail -40 gpurun_out/lab21_pytest.log
