#!/bin/bash
# N=4 re-check of the bench paths touched at the end of the session (Forest-Fire e2e leg with page-locked downloads; default line)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29560"
timeout 900 $TR bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/lab60_bench_n2.json 2> gpurun_out/lab60_bench_n2.err; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/lab60_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print('64M N=4', d['value'], d['ms_per_step'], (d.get('parity') or {}).get('mismatches'), (d.get('e2e') or {}).get('value'))
        print({k:(v['value'], v['ms_per_step'], (v.get('e2e') or {}).get('value')) for k,v in (d.get('extra') or {}).items() if isinstance(v, dict) and 'value' in v})
PY
tail -c 300 gpurun_out/lab60_bench_n2.err
