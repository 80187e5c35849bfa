#!/usr/bin/env python
"""CPU model of K4's candidate loops: how many loop trips a warp issues for a given thread->agent
mapping, and how many distinct 128-byte lines its candidate loads touch.  No GPU needed; the
numbers for the shipped mapping are checked against the executed-instruction counts of
profiles/r01_ncu_k4_packed_regions.txt.

usage: python tools/k4_lane_model.py [--agents 1000000] [--side 4000] [--seed 42] [--positions dump.npz]
"""
import argparse

import numpy as np

DISC = np.float32(10.0) / np.float32(1.5)


def build(n, side, seed, positions=None):
    if positions:
        d = np.load(positions)
        x, y = d["x"].astype(np.float32), d["y"].astype(np.float32)
        n = len(x)
    else:
        rng = np.random.default_rng(seed)
        x = (rng.random(n, dtype=np.float32) * np.float32(side)).astype(np.float32)
        y = (rng.random(n, dtype=np.float32) * np.float32(side)).astype(np.float32)
    max_x = int(np.ceil(np.float32(side) / DISC))
    dh = max_x + 1
    cx = np.floor(x / DISC).astype(np.int64)
    cy = np.floor(y / DISC).astype(np.int64)
    cell = cx * dh + cy
    order = np.argsort(cell, kind="stable")
    cx, cy, cell = cx[order], cy[order], cell[order]
    start = np.zeros(dh * dh + 1, np.int64)
    np.cumsum(np.bincount(cell, minlength=dh * dh), out=start[1:])
    lo = np.maximum(cy - 1, 0)
    hi = np.minimum(cy + 1, max_x - 1)
    s = np.zeros((3, n), np.int64)
    e = np.zeros((3, n), np.int64)
    for k, dx in enumerate((-1, 0, 1)):
        ci = cx + dx
        ok = (ci >= 0) & (ci <= max_x - 1) & (lo <= hi)
        cic = np.clip(ci, 0, max_x)
        s[k] = np.where(ok, start[cic * dh + lo], 0)
        e[k] = np.where(ok, start[cic * dh + np.maximum(hi, lo) + 1], 0)
    return s, e, cell


def warp_view(a, perm):
    n = len(perm) // 32 * 32
    return a[..., perm[:n]].reshape(*a.shape[:-1], n // 32, 32)


def report(name, s, e, perm, unroll=4):
    ln = warp_view(e - s, perm)                          # [3, warps, 32]
    main = (ln // unroll).max(axis=2)                    # trips of the unrolled body per slice
    tail = (ln % unroll).max(axis=2)
    slots = (main * unroll + tail).sum(axis=0)           # candidate slots issued per lane
    real = ln.sum(axis=0).mean()
    merged = ln.sum(axis=0)                              # the three slices as one sequence
    mslots = (merged // unroll).max(axis=1) * unroll + (merged % unroll).max(axis=1)
    # distinct 128 B lines (8 float4) touched by the first load of each slice
    first = warp_view(s, perm) // 8
    lines = np.array([[len(np.unique(first[k, w])) for w in range(0, first.shape[1], 97)] for k in range(3)])
    print(f"{name:34s} main/slice {main.mean():.2f}  tail/slice {tail.mean():.2f}  slots/lane {slots.mean():6.2f}"
          f"  (real {real:.1f}, utilisation {real / slots.mean():.0%})  merged-loop slots {mslots.mean():6.2f}"
          f"  lines/load {lines.mean():.1f}")
    return slots.mean()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=1_000_000)
    ap.add_argument("--side", type=float, default=4000.0)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--positions", help="npz with x, y (tools/dump_positions.py) instead of a uniform population")
    a = ap.parse_args()
    s, e, cell = build(a.agents, a.side, a.seed, a.positions)
    n = s.shape[1]
    ident = np.arange(n)
    total = (e - s).sum(axis=0)
    print(f"{n} agents, {a.side:g}^2, mean candidates per agent {total.mean():.2f}")
    base = report("shipped: buffer order, unroll 4", s, e, ident, 4)
    report("buffer order, unroll 2", s, e, ident, 2)
    report("buffer order, unroll 1", s, e, ident, 1)
    for tile in (256, 1024, 8192):
        perm = ident.copy()
        for t0 in range(0, n, tile):
            sl = slice(t0, min(t0 + tile, n))
            perm[sl] = t0 + np.argsort(total[sl], kind="stable")
        report(f"sorted by window length in tiles of {tile}", s, e, perm, 4)
    # whole cells sorted by window length inside a tile (lanes of one cell keep sharing addresses)
    for tile in (1024, 8192):
        perm = ident.copy()
        for t0 in range(0, n, tile):
            sl = slice(t0, min(t0 + tile, n))
            key = total[sl] * (1 << 32) + cell[sl]
            perm[sl] = t0 + np.argsort(key, kind="stable")
        report(f"cells sorted by length, tiles of {tile}", s, e, perm, 4)
    print(f"\nshipped mapping issues {base:.1f} candidate slots per lane.  profiles/r01_ncu_k4_packed_regions.txt (launch 5"
          f" of the bench): side main 4.45 + self main 2.24 trips x 4, side tails 2.00 (peeled first trip) + 3.89,"
          f" self tail 2.95 = {4 * (4.45 + 2.24) + 2.0 + 3.89 + 2.95:.1f}")

if __name__ == "__main__":
    main()
