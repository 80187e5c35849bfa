#!/usr/bin/env python
"""Step time of the exact-distance query (get_neighbors_within_distance) vs the relaxed one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krabmaga_b200 as kb  # noqa: E402

DISC = float(np.float32(10.0) / np.float32(1.5))


def main():
    for n in [int(a) for a in sys.argv[1:]] or [1_000_000]:
        w = float(np.sqrt(n / 0.0625))
        for exact in (0, 1):
            f = kb.Field2D(w, w, DISC, True, capacity=n)
            f.init_flockers(n, 42)
            f.lazy_update()
            p = kb.boids_params(radius=10.0, exact=exact, seed=42)
            f.run_boids(p, 5)
            p.step = 5
            f.profile(True)
            f.profile_read(reset=True)
            steps = 10
            ms = f.run_boids_timed(p, steps, 256 << 20)
            prof = f.profile_read(reset=True)
            print(f"n={n} exact={exact}: {1e3 * ms / steps:9.1f} us/step  "
                  f"{ {k: round(1e3 * v[0] / v[1], 1) for k, v in prof.items() if v[1]} }", flush=True)
            f.close()


if __name__ == "__main__":
    main()
