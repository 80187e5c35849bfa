#!/usr/bin/env python
"""Generates the golden fixtures in this directory from the oracle (oracle/ — the C++ restatement
of krABMaga's Field2D / DenseNumberGrid2D / Schedule / Flockers fixture).

    python tests/golden/make_golden.py          # rewrites *.npz / *.json next to this script

The reference itself cannot produce them: it is a Rust crate, this image has no cargo/rustc, and its
own tests hold no numeric vectors — only the count/membership known-answers that
tests/test_oracle_reference_kats.py ports.  These files pin the oracle (a later edit that changes
any bit fails tests/test_golden.py on CPU) and give the GPU tests a target that does not depend on
rebuilding the oracle.  Floats are stored as their u32 bit patterns.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_binding as ob  # noqa: E402
from parity_util import NORTH_STAR_DISC, random_agents  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def flockers_small():
    """2000 agents, 180x180 toroidal, disc 10/1.5, radius 10, relaxed + exact query, canonical
    in-bag order (ascending id) so that the summation order is part of the fixture."""
    n, w = 2000, 180.0
    agents = random_agents(n, w, w, seed=20261017)
    out = {k: (bits(v) if v.dtype == np.float32 else v) for k, v in agents.items()}
    for exact in (0, 1):
        m = ob.Flockers(w, w, n, NORTH_STAR_DISC, True, ob.boids_params(radius=10.0, exact=exact, seed=42),
                        canonical_order=True)
        m.preset(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
        m.init()
        done = 0
        for upto in (1, 5, 20):
            m.step(upto - done)
            done = upto
            x, y, dx, dy = m.agents()
            out[f"q{exact}_s{upto}"] = np.stack([bits(x), bits(y), bits(dx), bits(dy)])
        it = m.field1.iter_objects()
        cell = np.zeros(n, np.int32)
        cell[it["id"]] = it["cell"]
        out[f"q{exact}_cell20"] = cell
    np.savez_compressed(os.path.join(HERE, "flockers_small.npz"), **out)


def flockers_config1():
    """BASELINE config 1: 10,000 agents, 400x400, 200 steps, Philox init seed 42 (digest only)."""
    n, w = 10000, 400.0
    m = ob.Flockers(w, w, n, NORTH_STAR_DISC, True, ob.boids_params(radius=10.0, exact=0, seed=42),
                    canonical_order=True)
    m.init()
    rec = {"n": n, "w": w, "disc_bits": int(np.float32(NORTH_STAR_DISC).view(np.uint32)), "seed": 42,
           "steps": {}}
    done = 0
    for upto in (0, 1, 50, 200):
        m.step(upto - done)
        done = upto
        x, y, dx, dy = m.agents()
        rec["steps"][str(upto)] = {"sha256": digest(bits(x), bits(y), bits(dx), bits(dy)),
                                   "first8_x_bits": [int(v) for v in bits(x)[:8]],
                                   "first8_ldy_bits": [int(v) for v in bits(dy)[:8]]}
    with open(os.path.join(HERE, "flockers_config1.json"), "w") as f:
        json.dump(rec, f, indent=1)


def neighbors():
    """Neighbour id lists (reference order) for three geometries x both query kinds."""
    cases = {"fixture": (10.0, 10.0, 0.5, True, 150, [0.4, 1.0, 3.0, 10.0]),
             "northstar": (200.0, 200.0, NORTH_STAR_DISC, True, 2500, [10.0, 6.0, 25.0]),
             "nontoroidal": (120.0, 90.0, 4.5, False, 900, [10.0, 4.5])}
    out = {}
    for name, (w, h, d, t, n, radii) in cases.items():
        a = random_agents(n, w, h, seed=len(name) + n)
        rng = np.random.default_rng(n)
        qx = (rng.random(24, dtype=np.float32) * np.float32(w)).astype(np.float32)
        qy = (rng.random(24, dtype=np.float32) * np.float32(h)).astype(np.float32)
        qx[:4] = [0.0, w * 0.5, np.nextafter(np.float32(w), np.float32(0)), 0.01]
        qy[:4] = [0.0, h * 0.5, np.nextafter(np.float32(h), np.float32(0)), h - 0.01]
        f = ob.Field2D(w, h, d, t)
        f.set_object_locations(a["id"], a["x"], a["y"], a["ldx"], a["ldy"])
        f.lazy_update()
        out[f"{name}_geom"] = np.array([w, h, d, float(t)], np.float32)
        out[f"{name}_x"], out[f"{name}_y"] = bits(a["x"]), bits(a["y"])
        out[f"{name}_qx"], out[f"{name}_qy"] = bits(qx), bits(qy)
        it = f.iter_objects()
        cell = np.zeros(n, np.int32)
        cell[it["id"]] = it["cell"]
        out[f"{name}_cell"] = cell
        for ri, r in enumerate(radii):
            for mode in (0, 1):
                offs, ids = f.neighbors_batch(qx, qy, r, mode)
                out[f"{name}_r{ri}_m{mode}_offs"] = np.asarray(offs, np.int64)
                out[f"{name}_r{ri}_m{mode}_ids"] = np.asarray(ids, np.uint32)
        out[f"{name}_radii"] = np.array(radii, np.float32)
    np.savez_compressed(os.path.join(HERE, "neighbors.npz"), **out)


def forest_fire():
    w, h = 96, 64
    o = ob.ForestFire(w, h)
    o.init(0.6, 42)
    out = {"s0": o.dump().copy()}
    done = 0
    for upto in (1, 10, 40):
        o.step(upto - done)
        done = upto
        out[f"s{upto}"] = o.dump().copy()
    np.savez_compressed(os.path.join(HERE, "forest_fire_96x64.npz"), **out)


if __name__ == "__main__":
    flockers_small()
    flockers_config1()
    neighbors()
    forest_fire()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
