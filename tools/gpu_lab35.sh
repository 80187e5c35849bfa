#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blocks.py -x -q -m gpu > gpurun_out/lab35_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/lab35_pytest.log
tail -25 gpurun_out/lab35_pytest.log
python - <<'PY'
# timing of a block world next to the plain field and strips on one GPU (not a bench: what the host-orchestrated exchange costs)
import time, numpy as np, krabmaga_b200 as kb
from krabmaga_b200 import blocks
n=1_000_000; w=4000.0; DISC=float(np.float32(10.0)/np.float32(1.5))
f=kb.Field2D(w,w,DISC,True,capacity=n); f.init_flockers(n,42); init=f.download(unbuffered=True, with_cells=False); f.close()
p=kb.boids_params(radius=10.0, exact=0, seed=42)
for nbx,nby in ((1,1),(2,2),(4,2)):
    bw=blocks.BlockWorld(w,w,DISC,10.0,nbx,nby,[0],n,slack=2.0); bw.upload(init); p.step=0; bw.run_boids(p,5)
    t=time.perf_counter(); p.step=5; bw.run_boids(p,50); dt=time.perf_counter()-t
    print(f"blocks {nbx}x{nby} on one GPU, 1M agents: {1e6*dt/50:.1f} us/step (wall clock, host-orchestrated exchange)")
    bw.close()
PY
