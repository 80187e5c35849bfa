#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_custom_step.py -x -q -m gpu > gpurun_out/lab39_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/lab39_pytest.log
tail -5 gpurun_out/lab39_pytest.log
python - <<'PY' > gpurun_out/lab39_custom_timing.txt 2>&1
import sys, numpy as np
sys.path.insert(0, 'tests')
import krabmaga_b200 as kb
from krabmaga_b200 import _abi as abi
from custom_models import BIRD_PAIR, BIRD_FINISH
n=1_000_000; w=4000.0; DISC=float(np.float32(10.0)/np.float32(1.5))
p=kb.boids_params(radius=10.0, exact=0, seed=42)
def run(kind, exact):
    f=kb.Field2D(w,w,DISC,True,capacity=n); f.init_flockers(n,42); f.lazy_update()
    p.exact_query=exact
    if kind=="generic": f.set_kernel_variant(abi.KG_K4_GENERIC)
    def step(s):
        if kind=="custom": f.step_custom(BIRD_PAIR, BIRD_FINISH, [1,1,1,1,1,0.7], radius=10.0, exact=exact, seed=42, step=s)
        else:
            p.step=s; f.step_boids(p)
        f.lazy_update()
    for s in range(10): step(s)
    f.sync(); f.timer_start()
    for s in range(10,40): step(s)
    ms=f.timer_stop(); f.close()
    print(f"Bird::step, 1M agents, {'exact' if exact else 'relaxed'} query, {kind:8s}: {1e3*ms/30:7.1f} us/step", flush=True)
for exact in (0,1):
    for kind in ("packed","generic","custom"): run(kind, exact)
PY
cat gpurun_out/lab39_custom_timing.txt
