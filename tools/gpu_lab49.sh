#!/bin/bash
# T-step Forest Fire after the integer-pipe diet (table spread, interior fast path, uniform control flow): parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grid.py tests/test_gpu_gridstrips.py -q -m gpu -k "forest or strip or fire" > gpurun_out/lab49_pytest.log 2>&1; tail -5 gpurun_out/lab49_pytest.log
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --workload forest_fire --steps 400 --warmup 16 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$label', round(d['ms_per_step'], 4), 'ms/step', '%.3e' % d['value'], 'frac', round(d['roofline']['frac'], 3), 'launches', d['gpu_launches'])
"
}
run T8 A=1
run T4 KG_FF_FUSE=4
run T2 KG_FF_FUSE=2
for r in 64 160 256; do run T8_rows$r KG_FFT_ROWS=$r; done
for v in lut0 m8_6 m8_4; do run T8_$v KRABGPU_LIB=$PWD/gpurun_variants/libkrabgpu_$v.so; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forest_fire_u8_multi -s 3 -c 1 -o gpurun_out/lab49_ff_multi8 python bench.py --workload forest_fire --steps 40 --warmup 16 --no-cpu-baseline --no-e2e > gpurun_out/lab49_ncu.log 2>&1
