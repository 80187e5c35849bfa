#!/usr/bin/env python
"""Where does explore_parallel's wall time go?  (lab tool)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krabmaga_b200 as kb  # noqa: E402

DISC = float(np.float32(10.0) / np.float32(1.5))
n, reps, steps = 16384, 1024, 20
w = float(np.sqrt(n / 0.0625))
for trial in range(3):
    t = [time.perf_counter()]
    params = [kb.boids_params(radius=10.0, exact=0, seed=42 + r) for r in range(reps)]
    t.append(time.perf_counter())
    b = kb.FlockerBatch((w, w), n, reps, DISC, True, params)
    t.append(time.perf_counter())
    b.init(); b.sync()
    t.append(time.perf_counter())
    b.run(steps); b.sync()
    t.append(time.perf_counter())
    red = b.reduce()
    t.append(time.perf_counter())
    d = b.download()
    t.append(time.perf_counter())
    b.close()
    t.append(time.perf_counter())
    names = ["params", "create", "init", "run", "reduce", "download", "close"]
    print(trial, {k: round(1e3 * (t[i + 1] - t[i]), 1) for i, k in enumerate(names)}, flush=True)
for trial in range(2):
    t0 = time.perf_counter()
    rows = kb.explore_parallel(steps, 1, (w, w), n, DISC, {"seed": [42 + r for r in range(reps)]},
                               mode=kb.ExploreMode.Matched)
    print("explore_parallel ms", round(1e3 * (time.perf_counter() - t0), 1), len(rows), flush=True)
