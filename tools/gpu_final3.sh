#!/bin/bash
# clean re-take of the parts of gpu_final2.sh that the arbitrary-bytes test (K5's == 2 halo compare) had stopped
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/f3_pytest.log 2>&1; tail -3 gpurun_out/f3_pytest.log
{
echo "== memcheck: Forest Fire single grid + strips (fused passes of 8 / 4 / 2 steps, eight-row halos)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_grid.py tests/test_gpu_gridstrips.py -x -q -m gpu -k "two_steps or arbitrary or strips_equal_oracle or reupload" 2>&1 | tail -6
} > gpurun_out/f3_sanitizers.txt 2>&1
grep -E "==|passed|failed|ERROR SUMMARY" gpurun_out/f3_sanitizers.txt
