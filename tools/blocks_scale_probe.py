#!/usr/bin/env python
"""A block world over all visible GPUs, driven by this one process: wall-clock step time at the 64M-agent world
(not a bench line: the host-orchestrated exchange is what is being looked at).  usage: blocks_scale_probe.py [agents]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krabmaga_b200 as kb  # noqa: E402
from krabmaga_b200 import _abi as abi, blocks  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64_000_000
nd = abi.lib().kg_device_count()
DISC = float(np.float32(10.0) / np.float32(1.5))
w = float(np.sqrt(n / 0.0625))
f = kb.Field2D(w, w, DISC, True, capacity=n)
f.init_flockers(n, 42)
init = f.download(unbuffered=True, with_cells=False)
f.close()
p = kb.boids_params(radius=10.0, exact=0, seed=42)
shapes = {1: [(1, 1), (2, 2)], 2: [(2, 1)], 4: [(2, 2), (4, 1)], 8: [(4, 2), (8, 1)]}.get(nd, [(nd, 1)])
for nbx, nby in shapes:
    bw = blocks.BlockWorld(w, w, DISC, 10.0, nbx, nby, list(range(nd)), n, slack=1.5)
    bw.upload(init)
    p.step = 0
    bw.run_boids(p, 5)
    t = time.perf_counter()
    p.step = 5
    bw.run_boids(p, 30)
    dt = time.perf_counter() - t
    held = sum(a for a, _ in bw.counts())
    print(f"{n} agents, {nbx} x {nby} blocks on {nd} GPU(s): {1e3 * dt / 30:.3f} ms/step wall clock, "
          f"{n * 30 / dt:.3e} agent-steps/s, ghosts held {held - n}", flush=True)
    bw.close()
