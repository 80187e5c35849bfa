#!/bin/bash
# Round 2 record run (1 GPU): full GPU suite, default bench + reference arm, ncu launch list of the bench
# command, compute-sanitizer over the round's new kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/f1_pytest.log 2>&1; tail -3 gpurun_out/f1_pytest.log
timeout 900 python bench.py > gpurun_out/f1_bench_n1.json 2> gpurun_out/f1_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/f1_bench_ref_n1.json 2> gpurun_out/f1_bench_ref_n1.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f1_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-parity --no-scaling-ref --no-e2e > gpurun_out/f1_launches.log 2>&1
{
echo "== memcheck: object grid, dynamic population, K4 variants, queries"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_object_grid.py tests/test_gpu_life.py -x -q -m gpu -k "not schelling" 2>&1 | tail -8
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_field2d.py -x -q -m gpu -k "variant or queries_return or host_roundtrip or golden" 2>&1 | tail -8
echo "== racecheck + synccheck: tile kernel (shared memory, mbarrier), object grid"
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_field2d.py -x -q -m gpu -k "every_k4_variant" 2>&1 | tail -8
timeout 900 compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_gpu_field2d.py -x -q -m gpu -k "every_k4_variant" 2>&1 | tail -8
} > gpurun_out/f1_sanitizers.txt 2>&1
cat gpurun_out/f1_sanitizers.txt | grep -E "==|passed|failed|ERROR SUMMARY|RACECHECK SUMMARY"
python - <<'PY'
import json
for f in ("gpurun_out/f1_bench_n1.json","gpurun_out/f1_bench_ref_n1.json"):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), (d.get('parity') or {}).get('mismatches'), d['config']['workload'][:60])
            if d.get('extra'): print({k:(v['value'], (v.get('e2e') or {}).get('value')) for k,v in d['extra'].items()})
PY
