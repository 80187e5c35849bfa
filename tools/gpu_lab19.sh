#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_field2d.py tests/test_gpu_strips.py tests/test_gpu_batch.py tests/test_gpu_differential.py tests/test_gpu_life.py -q -m gpu > gpurun_out/lab19_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/lab19_pytest.log
tail -25 gpurun_out/lab19_pytest.log
{
timeout 300 python tools/k4_ab.py --agents 1000000 --variants 0,1 --exact 1 --steps 20
} > gpurun_out/lab19_ab.jsonl 2> gpurun_out/lab19_ab.err
cat gpurun_out/lab19_ab.jsonl; tail -3 gpurun_out/lab19_ab.err
python - <<'PY' > gpurun_out/lab19_dd.jsonl 2>&1
import json, numpy as np, krabmaga_b200 as kb
from krabmaga_b200 import _abi as abi
n=1_000_000; w=4000.0
for disc in (10/1.5, 10/2.5, 10/3.5):
    for variant in (abi.KG_K4_AUTO, abi.KG_K4_GENERIC):
        f=kb.Field2D(w,w,float(np.float32(disc)),True,capacity=n); f.set_kernel_variant(variant)
        f.init_flockers(n,42); f.lazy_update()
        p=kb.boids_params(radius=10.0, exact=1, seed=42); p.step=0
        f.run_boids(p,10); p.step=10
        ms=f.run_boids_timed(p,10,0)
        print(json.dumps({"disc":disc,"dd":int(10/disc),"variant":variant,"us_per_step":round(1e3*ms/10,1)}),flush=True)
        f.close()
PY
cat gpurun_out/lab19_dd.jsonl
