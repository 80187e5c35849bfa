#!/bin/bash
# T-steps-per-pass Forest Fire on bit planes: parity, then T / rows-per-tile / residency sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grid.py tests/test_gpu_gridstrips.py -q -m gpu -k "forest or strip or fire" > gpurun_out/lab45_pytest.log 2>&1; tail -15 gpurun_out/lab45_pytest.log
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --workload forest_fire --steps 200 --warmup 16 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$label', round(d['ms_per_step'], 4), 'ms/step', '%.3e' % d['value'], 'frac', round(d['roofline']['frac'], 3), 'launches', d['gpu_launches'])
"
}
run single KG_FF_FUSE=1
run T2 KG_FF_FUSE=2
run T4 KG_FF_FUSE=4
run T8 A=1
for r in 64 128 256; do run T8_rows$r KG_FFT_ROWS=$r; run T4_rows$r KG_FF_FUSE=4 KG_FFT_ROWS=$r; done
for v in m8_4 m8_6 m_8; do run T8_$v KRABGPU_LIB=$PWD/gpurun_variants/libkrabgpu_$v.so; run T4_$v KG_FF_FUSE=4 KRABGPU_LIB=$PWD/gpurun_variants/libkrabgpu_$v.so; done
