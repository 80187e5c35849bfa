#!/bin/bash
# GPU lab 7 (8 GPUs): the full bench line at N=8 (parity leg over 8 real strips, extras), per-kernel
# strip tables (KG_STRIP_PROF) and the round-1 exchange path for comparison
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/lab7_bench_n8.json 2> gpurun_out/lab7_bench_n8.err; echo "bench rc=$?"
KG_STRIP_PROF=1 timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-extra --no-parity --no-e2e > gpurun_out/lab7_n8_prof.json 2> gpurun_out/lab7_n8_prof.err
KG_STRIP_HALO=build KG_STRIP_PUSH=kernel timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-extra --no-parity --no-e2e > gpurun_out/lab7_n8_r01path.json 2> gpurun_out/lab7_n8_r01path.err
KG_STRIP_HALO=build KG_STRIP_PUSH=kernel KG_STRIP_PROF=1 timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-extra --no-parity --no-e2e > gpurun_out/lab7_n8_r01path_prof.json 2> gpurun_out/lab7_n8_r01path_prof.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/lab7_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['ms_per_step'], (d.get('parity') or {}).get('mismatches'), (d.get('e2e') or {}).get('value'), {k:(v['value'],v['ms_per_step']) for k,v in (d.get('extra') or {}).items()})
PY
grep -h "strip [037]\]" gpurun_out/lab7_n8_prof.err | sort | head -40
tail -c 600 gpurun_out/lab7_bench_n8.err
