#!/bin/bash
# by-position e2e entry (kg_field2d_step_boids_host_ordered): parity tests, then the default bench line's e2e (ordered vs keyed)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_field2d.py tests/test_golden.py -x -q -m gpu > gpurun_out/lab58_pytest.log 2>&1; tail -5 gpurun_out/lab58_pytest.log
timeout 600 python bench.py --no-extra --no-scaling-ref --no-parity --no-cpu-baseline > gpurun_out/lab58_bench.json 2> gpurun_out/lab58_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/lab58_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], json.dumps(d['e2e'], indent=1))
PY
tail -c 400 gpurun_out/lab58_bench.err
