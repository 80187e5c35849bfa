#!/bin/bash
# compute-sanitizer over the kernels of the round's second session
mkdir -p gpurun_out
{
echo "== memcheck: sparse + dense object grid, run-time closures, blocks, column-chunk / staged K4, wide exact windows, series"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_object_grid.py tests/test_gpu_closures.py -x -q -m gpu -k "not schelling and not large" 2>&1 | tail -6
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_blocks.py -x -q -m gpu -k "2-2 or 3-2 or wide or geometry" 2>&1 | tail -6
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_coltile.py tests/test_gpu_field2d.py -x -q -m gpu -k "coltile or every_k4_variant or series or (packed_exact and 21)" 2>&1 | tail -6
echo "== racecheck: column-chunk K4 (tables, staging, counting sort in shared memory), staged K4 (bulk copies + mbarrier)"
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_coltile.py -x -q -m gpu -k "bit_exact and (uniform or clustered or tiny)" 2>&1 | tail -6
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_field2d.py -x -q -m gpu -k "every_k4_variant and 10000" 2>&1 | tail -6
echo "== synccheck: the same"
timeout 900 compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_gpu_coltile.py tests/test_gpu_field2d.py -x -q -m gpu -k "(bit_exact and uniform) or (every_k4_variant and 10000)" 2>&1 | tail -6
} > gpurun_out/lab24_sanitizers.txt 2>&1
grep -E "==|passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error|error" gpurun_out/lab24_sanitizers.txt | head -40
