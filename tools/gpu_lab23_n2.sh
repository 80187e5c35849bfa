#!/bin/bash
# 2 GPUs: the bench line incl. the parity leg (strips, row strips, 2-D blocks over both GPUs)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 1200 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/lab23_bench_n2.json 2> gpurun_out/lab23_bench_n2.err; echo "rc=$?"
tail -c 1200 gpurun_out/lab23_bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/lab23_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d.get('parity'), (d.get('e2e') or {}).get('value'))
PY
