#!/usr/bin/env python
"""A/B timing of the K4 variants and the rebuild kernels on one GPU (not a bench: a lab tool).

    python tools/k4_ab.py [--agents 1000000] [--steps 40] [--variants 0,4] [--flush]

Prints, per variant, the CUDA-event time of the whole step and the isolated per-launch time of
every kernel kind (profiled pass), after letting the flock settle for --settle steps.  Variants
are the KG_K4_* values of include/krabgpu.h.  KRABGPU_LIB=<path> picks another build of the library.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DISC = float(np.float32(10.0) / np.float32(1.5))
DENSITY = 10000.0 / (400.0 * 400.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--settle", type=int, default=30)
    ap.add_argument("--variants", default="0,4")
    ap.add_argument("--flush", action="store_true")
    ap.add_argument("--exact", type=int, default=0)
    ap.add_argument("--check", action="store_true", help="compare every variant's one-step output bit for bit")
    args = ap.parse_args()
    import krabmaga_b200 as kb
    n = args.agents
    w = float(np.sqrt(n / DENSITY))
    params = kb.boids_params(radius=10.0, exact=args.exact, seed=42)
    f = kb.Field2D(w, w, DISC, True, capacity=n, device=0)
    f.init_flockers(n, 42)
    f.lazy_update()
    params.step = 0
    f.run_boids(params, args.settle)
    f.sync()
    flush = (256 << 20) if args.flush else 0
    ref = None
    for v in [int(x) for x in args.variants.split(",")]:
        f.set_kernel_variant(v)
        params.step = 1000
        f.run_boids(params, 3)
        params.step = 2000
        ms = f.run_boids_timed(params, args.steps, flush)
        f.profile(True)
        f.profile_read(reset=True)
        params.step = 3000
        f.run_boids_timed(params, args.steps, flush)
        prof = f.profile_read(reset=True)
        f.profile(False)
        kern = {k: round(1e3 * t / c, 2) for k, (t, c) in prof.items() if c and t > 0}
        line = {"lib": os.path.basename(kb._abi._SO), "variant": v, "agents": n, "exact": args.exact,
                "flush": bool(flush), "us_per_step": round(1e3 * ms / args.steps, 2),
                "agent_steps_per_s": n * args.steps / (ms * 1e-3), "kernels_us": kern}
        print(json.dumps(line), flush=True)
    if args.check:
        # one step from the same read buffer with every variant; compare by id
        outs = {}
        snap = f.download(with_cells=False)
        for v in [int(x) for x in args.variants.split(",")]:
            g = kb.Field2D(w, w, DISC, True, capacity=n, device=0)
            g.set_order(True)
            g.set_kernel_variant(v)
            g.set_object_locations(snap["id"], snap["x"], snap["y"], snap["ldx"], snap["ldy"])
            g.lazy_update()
            params.step = 77
            g.run_boids(params, 2)
            d = g.download(with_cells=False)
            o = np.argsort(d["id"])
            outs[v] = {k: d[k][o] for k in ("x", "y", "ldx", "ldy")}
            g.close()
        vs = list(outs)
        for v in vs[1:]:
            bad = sum(int((outs[v][k].view(np.uint32) != outs[vs[0]][k].view(np.uint32)).sum()) for k in outs[v])
            print(json.dumps({"check": f"variant {v} vs {vs[0]}", "mismatching_words": bad}), flush=True)
    f.close()


if __name__ == "__main__":
    main()
