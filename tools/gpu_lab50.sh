#!/bin/bash
# same-box A/B: committed T-step kernel (old) vs the restructured one with the multiply spread (lut0) and the table (default build)
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --workload forest_fire --steps 400 --warmup 16 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$label', round(d['ms_per_step'], 4), 'ms/step', '%.3e' % d['value'])
"
}
for i in 1 2; do
run old KRABGPU_LIB=$PWD/gpurun_variants/libkrabgpu_old.so
run lut0 KRABGPU_LIB=$PWD/gpurun_variants/libkrabgpu_lut0.so
run lut1 A=1
done
run single KG_FF_FUSE=1
