#!/bin/bash
# N=8: Forest Fire only, after the rows-per-tile wave model
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553"
timeout 300 $TR bench.py --gpus 8 --workload forest_fire --steps 1000 --warmup 16 > gpurun_out/lab53_bench_ff_n8.json 2> gpurun_out/lab53_bench_ff_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/lab53_bench_ff_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print('FF N=8', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], (d.get('e2e') or {}).get('value'))
PY
tail -c 200 gpurun_out/lab53_bench_ff_n8.err
