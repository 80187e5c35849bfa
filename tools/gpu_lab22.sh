#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/lab22_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/lab22_pytest.log
tail -8 gpurun_out/lab22_pytest.log
timeout 900 python bench.py > gpurun_out/lab22_bench_n1.json 2> gpurun_out/lab22_bench_n1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/lab22_bench_n1.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/lab22_bench_n1.json').read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["e2e"]["value"], l["parity"], l["roofline"]["frac"], l.get("gpu_launches"))
PY
