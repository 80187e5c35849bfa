#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
KG_FUZZ=1 timeout 900 python -m pytest tests/test_gpu_differential.py -x -q -m gpu 2>&1 | tail -3
done > gpurun_out/lab32_fuzz.log 2>&1
cat gpurun_out/lab32_fuzz.log | grep -E "passed|failed|Error|Falsifying" | head -20
