#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541"
timeout 900 $TR bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/lab30_bench_n4.json 2> gpurun_out/lab30_bench_n4.err; echo "bench rc=$?"
timeout 600 $TR bench.py --impl reference --gpus 4 --steps 20 --warmup 3 > gpurun_out/lab30_ref_n4.json 2> gpurun_out/lab30_ref_n4.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/lab30_bench_n4.json','gpurun_out/lab30_ref_n4.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['ms_per_step'], (d.get('parity') or {}).get('mismatches'), (d.get('e2e') or {}).get('value'), d['config'].get('workload','')[:80])
PY
tail -c 300 gpurun_out/lab30_bench_n4.err
