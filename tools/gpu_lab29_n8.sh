#!/bin/bash
# 8 GPUs, second session of round 2: the full bench line (parity: 8 strips, 8 row strips, 4 x 2 blocks over the 8 GPUs), strip tables
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/lab29_bench_n8.json 2> gpurun_out/lab29_bench_n8.err; echo "bench rc=$?"
KG_STRIP_PROF=1 timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-extra --no-parity --no-e2e > gpurun_out/lab29_n8_prof.json 2> gpurun_out/lab29_n8_prof.err
KG_STRIP_PREWAIT=0 KG_STRIP_FINISH=0 timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-extra --no-parity --no-e2e > gpurun_out/lab29_n8_old.json 2> gpurun_out/lab29_n8_old.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/lab29_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['ms_per_step'], (d.get('parity') or {}), (d.get('e2e') or {}).get('value'), {k:(v['value'],v['ms_per_step']) for k,v in (d.get('extra') or {}).items()})
PY
grep -h "strip [037]\]" gpurun_out/lab29_n8_prof.err | sort | grep -v "init\|halo_kernel" | head -30
tail -c 400 gpurun_out/lab29_bench_n8.err
