#!/bin/bash
# GPU lab 6 (2 GPUs): fused push (K4 stores into the peers' inboxes) — parity tests, then N=2 timing A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_strips.py tests/test_gpu_differential.py tests/test_gpu_field2d.py tests/test_gpu_full_size.py -x -q -m gpu > gpurun_out/lab6_pytest.log 2>&1; tail -3 gpurun_out/lab6_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
for mode in kernel fused; do
  KG_STRIP_PUSH=$mode timeout 600 $TR bench.py --gpus 2 --agents 16000000 --steps 20 --warmup 3 --no-extra --no-e2e > gpurun_out/lab6_n2_16m_$mode.json 2> gpurun_out/lab6_n2_16m_$mode.err
  KG_STRIP_PUSH=$mode KG_STRIP_PROF=1 timeout 600 $TR bench.py --gpus 2 --agents 16000000 --steps 20 --warmup 3 --no-extra --no-parity --no-e2e > gpurun_out/lab6_n2_16m_prof_$mode.json 2> gpurun_out/lab6_n2_16m_prof_$mode.err
done
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extra > gpurun_out/lab6_bench_n2.json 2> gpurun_out/lab6_bench_n2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/lab6_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['ms_per_step'], (d.get('parity') or {}).get('mismatches'), (d.get('e2e') or {}).get('value'))
PY
grep -h "strip 0\]" gpurun_out/lab6_n2_16m_prof_*.err | head -60
