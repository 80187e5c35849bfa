#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_strips.py tests/test_gpu_differential.py tests/test_gpu_full_size.py -x -q -m gpu > gpurun_out/lab28_pytest.log 2>&1; tail -3 gpurun_out/lab28_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523"
for ff in 0 1 0 1; do
  KG_STRIP_FINISH=$ff timeout 600 $TR bench.py --gpus 2 --agents 16000000 --steps 20 --warmup 3 --no-extra --no-e2e --no-parity > gpurun_out/lab28_ff$ff.json 2> gpurun_out/lab28_ff$ff.err
  python - <<PY
import json
for l in open('gpurun_out/lab28_ff$ff.json'):
    if l.startswith('{'):
        d=json.loads(l); print('fused_finish=$ff', d['value'], d['ms_per_step'])
PY
done
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extra > gpurun_out/lab28_bench_n2.json 2> gpurun_out/lab28_bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/lab28_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print('64M N=2', d['value'], d['ms_per_step'], (d.get('parity') or {}).get('mismatches'), (d.get('e2e') or {}).get('value'))
PY
tail -c 600 gpurun_out/lab28_bench_n2.err
