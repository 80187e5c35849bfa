#!/bin/bash
# last check of the session: whole GPU suite + default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/f5_pytest.log 2>&1; tail -2 gpurun_out/f5_pytest.log
timeout 600 python bench.py > gpurun_out/f5_bench_n1.json 2> gpurun_out/f5_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/f5_bench_n1.json"):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['e2e']['value'], d['parity']['mismatches'], d['roofline']['kernels']['step']['us_per_launch'])
        print({k:(v['value'], v.get('ms_per_step'), (v.get('e2e') or {}).get('value')) for k,v in d['extra'].items()})
PY
