#!/usr/bin/env python
"""Step time of G strips driven from ONE process (devices round-robin) vs the plain field:
separates the strip kernels' own cost (G=1: no neighbour) from the exchange (G=2 on two GPUs)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krabmaga_b200 as kb  # noqa: E402
from krabmaga_b200 import strips  # noqa: E402

DISC = float(np.float32(10.0) / np.float32(1.5))
per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
ndev = kb._abi.lib().kg_device_count()
for G in (1, 2):
    if G > ndev:
        break
    n = per_gpu * G
    w = float(np.sqrt(n / 0.0625))
    cap, hcap, mcap = strips.default_capacities(n, w, w, DISC, 10.0, G, slack=1.25)
    ss = [strips.StripField2D(w, w, DISC, 10.0, r, G, cap, hcap, mcap, device=r) for r in range(G)]
    for r, s in enumerate(ss):
        if G > 1:
            s.connect_local(ss[(r - 1) % G], ss[(r + 1) % G])
    for s in ss:
        s.init_flockers(n, 42)
    for s in ss:
        s.prepare()
    p = kb.boids_params(radius=10.0, exact=0, seed=42)
    for it in range(3):
        for s in ss:
            p.step = it * 10
            s.run_boids(p, 10)
        for s in ss:
            s.sync()
    for s in ss:
        s.timer_start()
    steps = 40
    for chunk in range(steps // 10):
        for s in ss:
            p.step = 30 + chunk * 10
            s.run_boids(p, 10)
    ms = max(s.timer_stop() for s in ss)
    print(f"G={G}: {1e3 * ms / steps:.1f} us/step for {per_gpu} agents per GPU", flush=True)
    for s in ss:
        s.close()
