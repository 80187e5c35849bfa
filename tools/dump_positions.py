#!/usr/bin/env python
"""Dump the positions of the default bench population after a few steps (input of
tools/k4_lane_model.py --positions) and time the step at those points.
usage (GPU box): python tools/dump_positions.py gpurun_out/positions 5 105"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import krabmaga_b200 as kb  # noqa: E402

DISC = float(np.float32(10.0) / np.float32(1.5))


def main():
    out = sys.argv[1]
    marks = [int(v) for v in sys.argv[2:]] or [5, 105]
    n, w = 1_000_000, 4000.0
    f = kb.Field2D(w, w, DISC, True, capacity=n)
    f.init_flockers(n, 42)
    f.lazy_update()
    p = kb.boids_params(radius=10.0, exact=0, seed=42)
    done = 0
    for m in marks:
        p.step = done
        f.run_boids(p, m - done)
        done = m
        d = f.download(with_cells=False)
        np.savez_compressed(f"{out}_step{m}.npz", x=d["x"], y=d["y"])
        p.step = done
        ms = f.run_boids_timed(p, 10, 256 << 20)
        done += 10
        print(f"after {m} steps: next 10 steps {ms / 10 * 1e3:.1f} us each (L2 flushed)")
    f.close()


if __name__ == "__main__":
    main()
