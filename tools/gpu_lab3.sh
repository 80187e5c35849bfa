#!/bin/bash
# GPU lab call 3 (round 2): packed K4 experiments (hoisted slice bounds, prefetch, software pipeline, occupancy)
set -x
mkdir -p gpurun_out
{
for lib in "" _pf _pf9 _base9 _pipe _r40; do
  KRABGPU_LIB=$PWD/krabmaga_b200/libkrabgpu$lib.so timeout 300 python tools/k4_ab.py --agents 1000000 --variants 0 --flush
done
for lib in "" _pf9 _pipe; do
  KRABGPU_LIB=$PWD/krabmaga_b200/libkrabgpu$lib.so timeout 300 python tools/k4_ab.py --agents 8000000 --variants 0 --steps 20
done
} > gpurun_out/lab3_ab.jsonl 2> gpurun_out/lab3_ab.err
cat gpurun_out/lab3_ab.jsonl
timeout 600 python -m pytest tests/test_gpu_field2d.py tests/test_golden.py -x -q -m gpu > gpurun_out/lab3_pytest.log 2>&1; tail -3 gpurun_out/lab3_pytest.log
