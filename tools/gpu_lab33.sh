#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/lab33_fail_*.log
for i in 1 2 3 4 5 6 7 8; do
KG_FUZZ=1 timeout 900 python -m pytest tests/test_gpu_differential.py -x -q -m gpu > gpurun_out/lab33_run.log 2>&1
if grep -q "failed" gpurun_out/lab33_run.log; then cp gpurun_out/lab33_run.log gpurun_out/lab33_fail_$i.log; fi
tail -1 gpurun_out/lab33_run.log
done
for f in gpurun_out/lab33_fail_*.log; do [ -f "$f" ] && grep -E "^(FAILED|E  )|Falsifying|case=|wts=|geom=|radius=|exact=" "$f" | head -40; done
