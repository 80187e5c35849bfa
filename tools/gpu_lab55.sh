#!/bin/bash
# look-back scan tile shapes at 8M agents (2.9M cells: the per-GPU size of the 64M world over 8 strips) and at 1M
for n in 8000000 1000000; do
for lib in "" lb256x32 lb512x16 lb512x32 lb1024x16; do
  if [ -n "$lib" ]; then export KRABGPU_LIB=$PWD/gpurun_variants/libkrabgpu_$lib.so; else unset KRABGPU_LIB; fi
  fl=""; [ $n = 1000000 ] && fl="--flush"
  timeout 200 python tools/k4_ab.py --agents $n --variants 0 --steps 40 $fl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print($n, d.get('lib'), d.get('us_per_step'), d.get('kernels_us'))"
done; done
