#!/bin/bash
# session 3 opening check: the whole GPU suite on HEAD
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/lab42_pytest.log 2>&1; tail -5 gpurun_out/lab42_pytest.log
