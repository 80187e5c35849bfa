#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grid.py tests/test_gpu_gridstrips.py -q -m gpu -k "forest or strip or fire" > gpurun_out/lab52_pytest.log 2>&1; tail -3 gpurun_out/lab52_pytest.log
python - <<'PY'
import os, subprocess, sys
CHILD = r"""
import sys, krabmaga_b200 as kb
w, h = int(sys.argv[1]), int(sys.argv[2])
g = kb.DenseNumberGrid2D(w, h)
g.init_forest_fire(0.6, 42)
g.run_stencil(16)
ms = g.run_stencil_timed(400)
print(ms / 400)
"""
for w, h in ((32768, 32768), (16384, 32768), (4096, 32768), (8192, 8192)):
    for fuse, rows in (("8", None), ("8", "107"), ("8", "160"), ("8", "200"), ("4", None), ("2", None)):
        env = dict(os.environ, KG_FF_FUSE=fuse)
        if rows: env["KG_FFT_ROWS"] = rows
        out = subprocess.run([sys.executable, "-c", CHILD, str(w), str(h)], env=env, capture_output=True, text=True)
        print(f"{w}x{h} T<={fuse} rows={rows or 'model'}: {out.stdout.strip() or out.stderr[-200:]} ms/step", flush=True)
PY
