#!/bin/bash
# two-steps-per-pass Forest Fire: parity, then timing against one step per pass (KG_FF_FUSE=0) and register budgets
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grid.py tests/test_gpu_gridstrips.py tests/test_gpu_full_size.py -q -m gpu -k "forest or strip or fire or grid" > gpurun_out/lab43_pytest.log 2>&1; tail -15 gpurun_out/lab43_pytest.log
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --workload forest_fire --steps 200 --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$label', round(d['ms_per_step'], 4), 'ms/step', '%.3e' % d['value'], 'frac', round(d['roofline']['frac'], 3), 'launches', d['gpu_launches'])
"
}
run single KG_FF_FUSE=0
run fused_minb8 A=1
for v in 7 6; do run fused_minb$v KRABGPU_LIB=$PWD/gpurun_variants/libkrabgpu_ff2_$v.so; done
run fused_minb8_rows128 KG_FF_ROWS=128
run fused_minb8_rows32 KG_FF_ROWS=32
