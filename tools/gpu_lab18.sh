#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_field2d.py tests/test_gpu_batch.py tests/test_gpu_life.py tests/test_gpu_strips.py -x -q -m gpu > gpurun_out/lab18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/lab18_pytest.log
tail -5 gpurun_out/lab18_pytest.log
{
for rep in 1 2; do
timeout 300 python tools/k4_ab.py --agents 1000000 --variants 0 --flush
KRABGPU_LIB=$PWD/krabmaga_b200/libkrabgpu_noearly.so timeout 300 python tools/k4_ab.py --agents 1000000 --variants 0 --flush
done
timeout 300 python tools/k4_ab.py --agents 8000000 --variants 0 --steps 20
KRABGPU_LIB=$PWD/krabmaga_b200/libkrabgpu_noearly.so timeout 300 python tools/k4_ab.py --agents 8000000 --variants 0 --steps 20
} > gpurun_out/lab18_ab.jsonl 2> gpurun_out/lab18_ab.err
cat gpurun_out/lab18_ab.jsonl; tail -3 gpurun_out/lab18_ab.err
