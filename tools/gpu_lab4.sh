#!/bin/bash
# GPU lab call 4 (round 2): rebuild experiments (warp-aggregated scatter, look-back scan shapes) + full gpu suite
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/lab4_pytest.log 2>&1; tail -3 gpurun_out/lab4_pytest.log
{
for lib in "" _lb128 _lb128x8 _lb512x8; do
  KRABGPU_LIB=$PWD/krabmaga_b200/libkrabgpu$lib.so timeout 300 python tools/k4_ab.py --agents 1000000 --variants 0 --flush
done
for lib in "" _lb128; do
  KRABGPU_LIB=$PWD/krabmaga_b200/libkrabgpu$lib.so timeout 300 python tools/k4_ab.py --agents 8000000 --variants 0 --steps 20
done
} > gpurun_out/lab4_ab.jsonl 2> gpurun_out/lab4_ab.err
cat gpurun_out/lab4_ab.jsonl
