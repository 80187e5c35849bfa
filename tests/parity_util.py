"""Shared helpers for the GPU-vs-oracle parity tests."""
import numpy as np

import oracle_binding as ob
from krabmaga_b200 import _abi as abi

NORTH_STAR_DISC = float(np.float32(10.0) / np.float32(1.5))  # 6.6666665f


def random_agents(n, w, h, seed, moving=True):
    rng = np.random.default_rng(seed)
    x = (rng.random(n, dtype=np.float32) * np.float32(w)).astype(np.float32)
    y = (rng.random(n, dtype=np.float32) * np.float32(h)).astype(np.float32)
    # keep strictly inside the world like toroidal_transform's usual output
    x = np.minimum(x, np.nextafter(np.float32(w), np.float32(0)))
    y = np.minimum(y, np.nextafter(np.float32(h), np.float32(0)))
    if moving:
        a = rng.random(n) * 2 * np.pi
        ldx, ldy = (0.7 * np.cos(a)).astype(np.float32), (0.7 * np.sin(a)).astype(np.float32)
    else:
        ldx = np.zeros(n, np.float32)
        ldy = np.zeros(n, np.float32)
    return dict(id=np.arange(n, dtype=np.uint32), x=x, y=y, ldx=ldx, ldy=ldy)


def both_params(**kw):
    return ob.boids_params(**kw), abi.boids_params(**kw)


def cells_by_id(d):
    out = np.zeros(len(d["id"]), np.int64)
    out[d["id"]] = d["cell"]
    return out


def bags(d):
    """dict cell -> sorted tuple of ids"""
    order = np.lexsort((d["id"], d["cell"]))
    cells, ids = d["cell"][order], d["id"][order]
    cut = np.flatnonzero(np.diff(cells)) + 1
    return {int(c[0]): tuple(int(v) for v in i)
            for c, i in zip(np.split(cells, cut), np.split(ids, cut)) if len(c)}


def csr_sets(offs, ids):
    return [tuple(sorted(int(v) for v in ids[offs[q]:offs[q + 1]])) for q in range(len(offs) - 1)]


def csr_lists(offs, ids):
    return [tuple(int(v) for v in ids[offs[q]:offs[q + 1]]) for q in range(len(offs) - 1)]


def by_id(d):
    """state arrays indexed by id from a download dict"""
    n = len(d["id"])
    out = {}
    for k in ("x", "y", "ldx", "ldy"):
        a = np.zeros(n, np.float32)
        a[d["id"]] = d[k]
        out[k] = a
    return out


def rel_err(a, b, scale):
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))) / scale)
