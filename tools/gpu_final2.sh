#!/bin/bash
# Round 2 record run, third session (1 GPU): full GPU suite, default bench + reference arm, ncu launch list of
# the bench command, compute-sanitizer over the Forest-Fire multi-step kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/f2_pytest.log 2>&1; tail -3 gpurun_out/f2_pytest.log
SECONDS=0; timeout 900 python bench.py > gpurun_out/f2_bench_n1.json 2> gpurun_out/f2_bench_n1.err; echo "bench rc=$? wall=${SECONDS}s"
timeout 600 python bench.py --impl reference > gpurun_out/f2_bench_ref_n1.json 2> gpurun_out/f2_bench_ref_n1.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f2_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-parity --no-scaling-ref --no-e2e > gpurun_out/f2_launches.log 2>&1
{
echo "== memcheck: Forest Fire single grid + strips (fused passes of 8 / 4 / 2 steps, eight-row halos)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_grid.py tests/test_gpu_gridstrips.py -x -q -m gpu -k "two_steps or arbitrary or strips_equal_oracle or reupload" 2>&1 | tail -8
echo "== racecheck: shared-memory row ring of the multi-step kernel"
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_grid.py -x -q -m gpu -k "two_steps and (65 or 130 or 2-16)" 2>&1 | tail -8
echo "== synccheck"
timeout 900 compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_gpu_gridstrips.py -x -q -m gpu -k "two_steps and (12-32 or 131)" 2>&1 | tail -8
} > gpurun_out/f2_sanitizers.txt 2>&1
cat gpurun_out/f2_sanitizers.txt | grep -E "==|passed|failed|ERROR SUMMARY|RACECHECK SUMMARY"
python - <<'PY'
import json
for f in ("gpurun_out/f2_bench_n1.json","gpurun_out/f2_bench_ref_n1.json"):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), (d.get('parity') or {}).get('mismatches'), d['config']['workload'][:60])
            if d.get('extra'): print({k:(v['value'], v.get('ms_per_step'), (v.get('e2e') or {}).get('value')) for k,v in d['extra'].items()})
PY
