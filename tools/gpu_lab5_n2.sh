#!/bin/bash
# GPU lab 5 (2 GPUs): the new bench flow at N=2 (parity leg over real NVLink, extras), and strips
# halo paths A/B with per-kernel tables (KG_STRIP_PROF)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/lab5_bench_n2.json 2> gpurun_out/lab5_bench_n2.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/lab5_bench_n2.err
for mode in build fold; do
  KG_STRIP_HALO=$mode timeout 600 $TR bench.py --gpus 2 --agents 16000000 --steps 20 --warmup 3 --no-extra --no-parity --no-e2e > gpurun_out/lab5_n2_16m_$mode.json 2> gpurun_out/lab5_n2_16m_$mode.err
  KG_STRIP_HALO=$mode KG_STRIP_PROF=1 timeout 600 $TR bench.py --gpus 2 --agents 16000000 --steps 20 --warmup 3 --no-extra --no-parity --no-e2e > gpurun_out/lab5_n2_16m_prof_$mode.json 2> gpurun_out/lab5_n2_16m_prof_$mode.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/lab5_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['ms_per_step'], d.get('parity'))
PY
grep -h "strip 0\]" gpurun_out/lab5_n2_16m_prof_*.err | head -60
