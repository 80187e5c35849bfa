#!/bin/bash
# ncu --set full of the eight-steps-per-pass Forest-Fire kernel (32768^2) + launch list of the forest_fire bench command
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:forest_fire_u8_multi -s 3 -c 1 -o gpurun_out/lab46_ff_multi8 python bench.py --workload forest_fire --steps 40 --warmup 16 --no-cpu-baseline --no-e2e > gpurun_out/lab46_ncu.log 2>&1
tail -2 gpurun_out/lab46_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/lab46_ff_launches.csv python bench.py --workload forest_fire --steps 40 --warmup 16 --no-cpu-baseline --no-e2e > gpurun_out/lab46_launches.log 2>&1
tail -2 gpurun_out/lab46_launches.log
timeout 300 python bench.py --workload forest_fire --steps 1000 --warmup 16 > gpurun_out/lab46_bench_ff_n1.json 2> gpurun_out/lab46_bench_ff_n1.err; tail -c 600 gpurun_out/lab46_bench_ff_n1.json
