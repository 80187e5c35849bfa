#!/bin/bash
# scatter with run-length aggregated rank allocation (KG_SCATTER_RUNS=1): parity with the variant library, then timing
mkdir -p gpurun_out
KRABGPU_LIB=$PWD/gpurun_variants/libkrabgpu_runs.so timeout 600 python -m pytest tests/test_gpu_field2d.py tests/test_gpu_life.py tests/test_golden.py -x -q -m gpu > gpurun_out/lab57_pytest.log 2>&1; tail -3 gpurun_out/lab57_pytest.log
for n in 1000000 8000000; do
for lib in "" runs runs1 "" runs; do
  if [ -n "$lib" ]; then export KRABGPU_LIB=$PWD/gpurun_variants/libkrabgpu_$lib.so; else unset KRABGPU_LIB; fi
  fl=""; [ $n = 1000000 ] && fl="--flush"
  timeout 200 python tools/k4_ab.py --agents $n --variants 0 --steps 40 $fl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print($n, d.get('lib'), d.get('us_per_step'), d.get('kernels_us'))"
done; done
