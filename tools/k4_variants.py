#!/usr/bin/env python
"""A/B timing of the K4 variants (same state, same steps): per-kernel CUDA-event times.
usage: python tools/k4_variants.py [n_agents ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import krabmaga_b200 as kb  # noqa: E402
from krabmaga_b200 import _abi as abi  # noqa: E402

DISC = float(np.float32(10.0) / np.float32(1.5))
NAMES = {abi.KG_K4_AUTO: "packed", abi.KG_K4_GENERIC: "generic", abi.KG_K4_FAST_SCALAR: "scalar",
         abi.KG_K4_PACKED_BY_ID: "packed_by_id", abi.KG_K4_TILED: "tiled",
         abi.KG_K4_COLTILE: "coltile", abi.KG_K4_STAGED: "staged"}


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1_000_000]
    for n in sizes:
        w = float(np.sqrt(n / 0.0625))
        f = kb.Field2D(w, w, DISC, True, capacity=n)
        f.init_flockers(n, 42)
        f.lazy_update()
        p = kb.boids_params(radius=10.0, exact=0, seed=42)
        f.run_boids(p, 10)
        for variant in (abi.KG_K4_AUTO, abi.KG_K4_TILED, abi.KG_K4_AUTO, abi.KG_K4_TILED):
            f.set_kernel_variant(variant)
            p.step = 10
            f.run_boids(p, 3)
            f.profile(True)
            f.profile_read(reset=True)
            steps = 20
            ms = f.run_boids_timed(p, steps, 256 << 20)
            prof = f.profile_read(reset=True)
            f.profile(False)
            line = {k: round(1e3 * v[0] / v[1], 1) for k, v in prof.items() if v[1]}
            print(f"n={n} {NAMES[variant]:>13}: {1e3 * ms / steps:8.1f} us/step  "
                  f"{n * steps / (ms * 1e-3):.3e} agent-steps/s  per-kernel us {line}", flush=True)
        f.close()


if __name__ == "__main__":
    main()
