#!/bin/bash
# K4: four-candidate body written out twice (KG_K4_ROT2) vs the default build, same box, bit-exactness vs the generic kernel
for i in 1 2; do
for lib in "" gpurun_variants/libkrabgpu_rot2.so gpurun_variants/libkrabgpu_rot2_m9.so; do
  if [ -n "$lib" ]; then export KRABGPU_LIB=$PWD/$lib; else unset KRABGPU_LIB; fi
  timeout 200 python tools/k4_ab.py --variants 0 --steps 100 --flush | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d.get('lib'), d.get('us_per_step'), d.get('kernels_us'))"
done; done
export KRABGPU_LIB=$PWD/gpurun_variants/libkrabgpu_rot2.so
timeout 200 python tools/k4_ab.py --variants 1,0 --steps 5 --check | tail -1
timeout 200 python tools/k4_ab.py --agents 8000000 --variants 0 --steps 20 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('8M', d.get('lib'), d.get('us_per_step'), d.get('kernels_us'))"
unset KRABGPU_LIB
timeout 200 python tools/k4_ab.py --agents 8000000 --variants 0 --steps 20 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('8M', d.get('lib'), d.get('us_per_step'), d.get('kernels_us'))"
