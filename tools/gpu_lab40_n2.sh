#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
timeout 1200 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/lab40_bench_n2.json 2> gpurun_out/lab40_bench_n2.err; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/lab40_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], (d.get('parity') or {}).get('mismatches'), d['extra'].get('block_world'))
PY
tail -c 300 gpurun_out/lab40_bench_n2.err
timeout 600 python -m pytest tests/test_gpu_object_grid.py -q -m gpu 2>&1 | tail -2
