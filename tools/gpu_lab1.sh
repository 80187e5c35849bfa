#!/bin/bash
# GPU lab call 1 (round 2): parity of the new K4 tile kernel + epilogue, A/B timings, ncu captures.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/lab1_smi.txt
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/lab1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/lab1_pytest.log
tail -5 gpurun_out/lab1_pytest.log
{
for fl in "" "--flush"; do
  timeout 300 python tools/k4_ab.py --agents 1000000 --variants 0,4 --check $fl
  KRABGPU_LIB=$PWD/krabmaga_b200/libkrabgpu_mt.so timeout 300 python tools/k4_ab.py --agents 1000000 --variants 0 $fl
done
for t in 150 200 260 300; do KG_TILE_TARGET=$t timeout 300 python tools/k4_ab.py --agents 1000000 --variants 4 --flush; done
timeout 300 python tools/k4_ab.py --agents 8000000 --variants 0,4 --steps 20
KRABGPU_LIB=$PWD/krabmaga_b200/libkrabgpu_mt.so timeout 300 python tools/k4_ab.py --agents 8000000 --variants 0 --steps 20
} > gpurun_out/lab1_ab.jsonl 2> gpurun_out/lab1_ab.err
cat gpurun_out/lab1_ab.jsonl
# launch list and one full capture of the tile kernel and the packed kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/lab1_launches.csv python tools/k4_ab.py --agents 1000000 --variants 4 --steps 10 --settle 30 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_boids_tile -s 40 -c 2 -o gpurun_out/lab1_tile python tools/k4_ab.py --agents 1000000 --variants 4 --steps 5 --settle 30 > gpurun_out/lab1_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_boids_packed -s 40 -c 2 -o gpurun_out/lab1_packed python tools/k4_ab.py --agents 1000000 --variants 0 --steps 5 --settle 30 >> gpurun_out/lab1_ncu.log 2>&1
ls -la gpurun_out
# the new bench flow (extras + parity leg + blocks), N=1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/lab1_bench_n1.json 2> gpurun_out/lab1_bench_n1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/lab1_bench_n1.err
