#!/usr/bin/env python
"""Run a few steps of one K4 variant (for ncu captures): python tools/run_variant.py VARIANT [n] [exact]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krabmaga_b200 as kb  # noqa: E402

variant = int(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
exact = int(sys.argv[3]) if len(sys.argv) > 3 else 0
disc = float(np.float32(10.0) / np.float32(1.5))
w = float(np.sqrt(n / 0.0625))
f = kb.Field2D(w, w, disc, True, capacity=n)
f.init_flockers(n, 42)
f.lazy_update()
f.set_kernel_variant(variant)
p = kb.boids_params(radius=10.0, exact=exact, seed=42)
f.run_boids(p, 8)
f.sync()
