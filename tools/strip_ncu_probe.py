#!/usr/bin/env python
"""Two strips of one world on ONE GPU (in-process StripWorld), for an ncu capture of strip_step_kernel next to the
plain field's packed K4 at the same number of agents per launch.  usage: python tools/strip_ncu_probe.py [agents_per_strip]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krabmaga_b200 as kb  # noqa: E402
from krabmaga_b200 import strips  # noqa: E402

DISC = float(np.float32(10.0) / np.float32(1.5))
per = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
n = 2 * per
w = float(np.sqrt(n / 0.0625))
p = kb.boids_params(radius=10.0, exact=0, seed=42)
world = strips.StripWorld(w, w, DISC, 10.0, [0, 0], n, slack=2.0)
world.init_flockers(n, 42)
p.step = 0
world.run_boids(p, 12)
for s in world.strips:
    s.sync()
world.close()
f = kb.Field2D(float(np.sqrt(per / 0.0625)), float(np.sqrt(per / 0.0625)), DISC, True, capacity=per)
f.init_flockers(per, 42)
f.lazy_update()
p.step = 0
f.run_boids(p, 12)
f.sync()
f.close()
print("done")
