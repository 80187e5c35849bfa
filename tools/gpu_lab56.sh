#!/bin/bash
# scatter: log entries per thread (1 / 2 default / 3 / 4) at 1M (L2 flushed between steps) and 8M agents
for n in 1000000 8000000; do
for lib in "" si1 si3 si4; do
  if [ -n "$lib" ]; then export KRABGPU_LIB=$PWD/gpurun_variants/libkrabgpu_$lib.so; else unset KRABGPU_LIB; fi
  fl=""; [ $n = 1000000 ] && fl="--flush"
  timeout 200 python tools/k4_ab.py --agents $n --variants 0 --steps 40 $fl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print($n, d.get('lib'), d.get('us_per_step'), d.get('kernels_us'))"
done; done
