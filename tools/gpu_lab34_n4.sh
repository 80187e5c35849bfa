#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/lab34_topo.txt 2>&1; head -20 gpurun_out/lab34_topo.txt
nproc; python -c "import os; print(sorted(os.sched_getaffinity(0)))"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543"
for aff in 0 1; do
KG_BENCH_AFFINITY=$aff timeout 900 $TR bench.py --gpus 4 --steps 20 --warmup 3 --no-extra --no-parity > gpurun_out/lab34_n4_aff$aff.json 2> gpurun_out/lab34_n4_aff$aff.err
python - <<PY
import json
for l in open('gpurun_out/lab34_n4_aff$aff.json'):
    if l.startswith('{'):
        d=json.loads(l); print('affinity=$aff', d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d.get('host_affinity'))
PY
done
