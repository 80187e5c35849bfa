#!/bin/bash
# 2 GPUs: strips with the single-thread pre-wait (publish kernel waits for the incoming flags) vs every append block waiting
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_strips.py tests/test_gpu_differential.py -x -q -m gpu > gpurun_out/lab25_pytest.log 2>&1; tail -3 gpurun_out/lab25_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
for pw in 0 1 0 1; do
  KG_STRIP_PREWAIT=$pw timeout 600 $TR bench.py --gpus 2 --agents 16000000 --steps 20 --warmup 3 --no-extra --no-e2e --no-parity > gpurun_out/lab25_n2_pw$pw.json 2> gpurun_out/lab25_n2_pw$pw.err
  python - <<PY
import json
for l in open('gpurun_out/lab25_n2_pw$pw.json'):
    if l.startswith('{'):
        d=json.loads(l); print('prewait=$pw', d['value'], d['ms_per_step'])
PY
done
for pw in 0 1; do
  KG_STRIP_PREWAIT=$pw KG_STRIP_PROF=1 timeout 600 $TR bench.py --gpus 2 --agents 16000000 --steps 20 --warmup 3 --no-extra --no-parity --no-e2e > /dev/null 2> gpurun_out/lab25_n2_prof_pw$pw.err
  echo "prewait=$pw"; grep -h "strip 0\]" gpurun_out/lab25_n2_prof_pw$pw.err | grep -v "init\|halo_kernel"
done
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extra --no-e2e > gpurun_out/lab25_bench_n2.json 2> gpurun_out/lab25_bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/lab25_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print('64M N=2', d['value'], d['ms_per_step'], (d.get('parity') or {}).get('mismatches'))
PY
