#!/bin/bash
# packed K4 register budget: 10 (default, 48 regs) / 9 (56) / 8 (64) resident blocks per SM, same box, twice
for i in 1 2; do
for lib in "" m9 m8; do
  if [ -n "$lib" ]; then export KRABGPU_LIB=$PWD/gpurun_variants/libkrabgpu_$lib.so; else unset KRABGPU_LIB; fi
  timeout 200 python tools/k4_ab.py --variants 0 --steps 100 --flush | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('1M', d.get('lib'), d.get('us_per_step'), d.get('kernels_us'))"
done; done
for lib in "" m9 m8; do
  if [ -n "$lib" ]; then export KRABGPU_LIB=$PWD/gpurun_variants/libkrabgpu_$lib.so; else unset KRABGPU_LIB; fi
  timeout 200 python tools/k4_ab.py --agents 8000000 --variants 0 --steps 20 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('8M', d.get('lib'), d.get('us_per_step'), d.get('kernels_us'))"
done
