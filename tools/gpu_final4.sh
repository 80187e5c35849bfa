#!/bin/bash
# re-take of the default bench line and the reference arm after the e2e changes (by-position Flockers entry, page-locked
# Forest-Fire download), the ordered-entry tests, and the launch list of the bench command
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_field2d.py tests/test_gpu_gridstrips.py -x -q -m gpu > gpurun_out/f4_pytest.log 2>&1; tail -2 gpurun_out/f4_pytest.log
SECONDS=0; timeout 900 python bench.py > gpurun_out/f4_bench_n1.json 2> gpurun_out/f4_bench_n1.err; echo "bench rc=$? wall=${SECONDS}s"
timeout 600 python bench.py --impl reference > gpurun_out/f4_bench_ref_n1.json 2> gpurun_out/f4_bench_ref_n1.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/f4_bench_n1.json","gpurun_out/f4_bench_ref_n1.json"):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), ((d.get('e2e') or {}).get('keyed') or {}).get('value'), (d.get('parity') or {}).get('mismatches'))
            if d.get('extra'): print({k:(v['value'], v.get('ms_per_step'), (v.get('e2e') or {}).get('value')) for k,v in d['extra'].items()})
PY
