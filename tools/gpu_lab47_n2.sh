#!/bin/bash
# N=2: grid strips on two real GPUs (fused passes, eight-row halos over NVLink), then the default bench line at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gridstrips.py tests/test_gpu_grid.py -x -q -m gpu > gpurun_out/lab47_pytest.log 2>&1; tail -3 gpurun_out/lab47_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547"
timeout 600 $TR bench.py --gpus 2 --workload forest_fire --steps 1000 --warmup 16 > gpurun_out/lab47_bench_ff_n2.json 2> gpurun_out/lab47_bench_ff_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/lab47_bench_ff_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print('FF N=2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], (d.get('e2e') or {}).get('value'))
PY
tail -c 300 gpurun_out/lab47_bench_ff_n2.err
timeout 1200 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/lab47_bench_n2.json 2> gpurun_out/lab47_bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/lab47_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print('64M N=2', d['value'], d['ms_per_step'], d.get('parity'), (d.get('e2e') or {}).get('value'))
        print({k:(v['value'], v['ms_per_step']) for k,v in (d.get('extra') or {}).items() if isinstance(v, dict) and 'value' in v})
PY
tail -c 300 gpurun_out/lab47_bench_n2.err
