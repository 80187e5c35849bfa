#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/blocks_scale_probe.py 64000000 > gpurun_out/lab37_blocks_n8.txt 2>&1; tail -4 gpurun_out/lab37_blocks_n8.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-extra --no-e2e > gpurun_out/lab37_bench_n8.json 2> gpurun_out/lab37_bench_n8.err; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/lab37_bench_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d.get('parity'))
PY
tail -c 300 gpurun_out/lab37_bench_n8.err
