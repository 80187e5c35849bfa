#!/bin/bash
# GPU lab call 2 (round 2): FP32 pipe probe, tile kernel staging variants, ncu of the tile kernel.
set -x
mkdir -p gpurun_out
./tools/probe/pipe_probe > gpurun_out/lab2_pipe_probe.jsonl 2>&1
cat gpurun_out/lab2_pipe_probe.jsonl
{
timeout 300 python tools/k4_ab.py --agents 1000000 --variants 0,4
KG_TILE_STAGE=1 timeout 300 python tools/k4_ab.py --agents 1000000 --variants 4
} > gpurun_out/lab2_ab.jsonl 2> gpurun_out/lab2_ab.err
cat gpurun_out/lab2_ab.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_boids_tile -s 4 -c 1 -o gpurun_out/lab2_tile python tools/k4_ab.py --agents 1000000 --variants 4 --steps 5 --settle 30 > gpurun_out/lab2_ncu.log 2>&1
KG_TILE_STAGE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_boids_tile -s 4 -c 1 -o gpurun_out/lab2_tile_ldg python tools/k4_ab.py --agents 1000000 --variants 4 --steps 5 --settle 30 >> gpurun_out/lab2_ncu.log 2>&1
timeout 600 python -m pytest tests/test_gpu_strips.py tests/test_gpu_grid.py -x -q -m gpu > gpurun_out/lab2_pytest.log 2>&1; tail -3 gpurun_out/lab2_pytest.log
ls -la gpurun_out
