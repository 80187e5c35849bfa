#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_field2d.py -q -m gpu -k "series" > gpurun_out/lab20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/lab20_pytest.log
tail -25 gpurun_out/lab20_pytest.log
