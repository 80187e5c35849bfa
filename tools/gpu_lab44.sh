#!/bin/bash
# two-steps-per-pass Forest Fire with the cp.async row ring: parity, then ring depth / residency sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grid.py tests/test_gpu_gridstrips.py -q -m gpu -k "forest or strip or fire" > gpurun_out/lab44_pytest.log 2>&1; tail -5 gpurun_out/lab44_pytest.log
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
run ring4_m8 A=1
for v in ring0_m7 ring2 ring3 ring6 ring8 ring4_m9 ring4_m10; do run $v KRABGPU_LIB=$PWD/gpurun_variants/libkrabgpu_$v.so; done
run ring4_m8_rows128 KG_FF_ROWS=128
run ring4_m8_rows256 KG_FF_ROWS=256
