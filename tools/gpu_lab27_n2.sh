#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
for lib in libkrabgpu.so libkrabgpu_lb10.so libkrabgpu.so libkrabgpu_lb10.so; do
  KRABGPU_LIB=$PWD/krabmaga_b200/$lib timeout 600 $TR bench.py --gpus 2 --agents 16000000 --steps 20 --warmup 3 --no-extra --no-e2e --no-parity > gpurun_out/lab27_$lib.json 2> gpurun_out/lab27_$lib.err
  python - <<PY
import json
for l in open('gpurun_out/lab27_$lib.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$lib', d['value'], d['ms_per_step'])
PY
done
