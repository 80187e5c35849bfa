"""The three BASELINE.json configurations that do not fit an oracle run, at their FULL sizes on one
GPU, through properties that do not depend on the size (conservation, sortedness, histogram =
download, |last_d| = JUMP, replica isolation, fire-front speed) plus oracle spot checks on a
slab of the world.  KG_SKIP_FULL=1 skips the module (≈6 GB of host memory, ≈1 min)."""
import os

import numpy as np
import pytest

import krabmaga_b200 as kb
import oracle_binding as ob
from krabmaga_b200 import _abi as abi
from parity_util import NORTH_STAR_DISC, both_params, csr_sets

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("KG_SKIP_FULL", "") not in ("", "0"), reason="KG_SKIP_FULL")]


def jump_or_still(ldx, ldy):
    """|last_d| == JUMP (bird.rs:139-143) — or 0 for an agent that has never had a neighbour"""
    norm = np.hypot(ldx, ldy)
    still = norm == 0
    return bool((still | (np.abs(norm - np.float32(0.7)) < 1e-5)).all()) and still.mean() < 1e-6


def test_config3_64m_agents_in_one_world():
    """config 3: 64 000 000 agents, 32000 x 32000, toroidal, radius 10 (here on one GPU; the strip
    decomposition of the same world is compared with one GPU in test_gpu_strips.py)"""
    n, w = 64_000_000, 32000.0
    _, gp = both_params(exact=0, seed=42)
    f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=n)
    f.init_flockers(n, 42)
    f.lazy_update()
    f.run_boids(gp, 3)
    assert f.num_objects() == n
    d = f.download()
    counts = f.cell_counts()
    assert len(d["id"]) == n and int(counts.sum(dtype=np.int64)) == n
    seen = np.zeros(n, bool)
    seen[d["id"]] = True                                            # ids < n, or this raises
    assert seen.all()                                               # n entries, all ids present: no twin
    del seen
    assert (np.diff(d["cell"]) >= 0).all()                          # sorted by flat cell index
    assert (np.bincount(d["cell"], minlength=len(counts)) == counts).all()
    for k in ("x", "y"):
        assert d[k].min() >= 0 and d[k].max() <= w
    assert jump_or_still(d["ldx"], d["ldy"])
    # neighbour sets against the oracle on a slab of five cell columns in the middle of the world
    # (the window of an agent of the middle column lies inside the slab)
    col = d["cell"] // f.dh
    cx = f.dw // 2
    slab = np.flatnonzero((col >= cx - 2) & (col <= cx + 2))
    mid = np.flatnonzero(col == cx)
    del col
    o = ob.Field2D(w, w, NORTH_STAR_DISC, True)
    o.set_object_locations(d["id"][slab], d["x"][slab], d["y"][slab], d["ldx"][slab], d["ldy"][slab])
    o.lazy_update()
    q = mid[np.random.default_rng(3).choice(len(mid), 1500, replace=False)]
    qx, qy = d["x"][q].copy(), d["y"][q].copy()
    offs, ids = f.neighbors_batch(np.stack([qx, qy], 1), 10.0, exact=False)
    ooffs, oids = o.neighbors_batch(qx, qy, 10.0, 0)
    assert csr_sets(offs, ids) == csr_sets(ooffs, oids)
    offs, ids = f.neighbors_batch(np.stack([qx, qy], 1), 10.0, exact=True)
    ooffs, oids = o.neighbors_batch(qx, qy, 10.0, 1)
    assert csr_sets(offs, ids) == csr_sets(ooffs, oids)
    f.close()


def test_config4_forest_fire_32768_squared():
    """config 4: 2^30 cells of u8.  None never changes, states only grow, the front advances one
    row per step, and the burning band equals the oracle run on a slab cut out of the grid."""
    w = h = 32768
    steps = 24
    g = kb.DenseNumberGrid2D(w, h)
    g.init_forest_fire(0.6, 42)
    a = g.download().copy()
    g.run_stencil(steps)
    b = g.download()
    none_a = a == 0xFF
    assert 0.39 < none_a.mean() < 0.41                              # density 0.6 of live cells
    assert np.array_equal(none_a, b == 0xFF)
    assert (b >= a).all()                                           # GREEN < BURNING < BURNED
    assert ((b[steps + 1:] == 1) | none_a[steps + 1:]).all()        # nothing beyond the front
    assert ((b[0] == 3) | none_a[0]).all()
    assert (b[:steps + 1] == 2).any()
    # a slab of the ignition side against the oracle run on the same initial cells.  Rows beyond
    # the slab stay GREEN/None for `steps` steps (no influence); beyond its last column the grid
    # has fire the oracle does not see, and that difference travels one cell per step
    top, wide = 40, 2048
    o = ob.ForestFire(top, wide)
    o.load(a[:top, :wide].copy())
    o.step(steps)
    want = o.dump()
    assert (b[:top, :wide - steps] == want[:, :wide - steps]).all()
    assert (want == 2).any() and (want == 3).any()
    g.close()


def test_config5_sweep_of_4096_replicas():
    """config 5: 4096 replicas of 16 384 agents in 512 x 512 as ONE batch: agents conserved per
    replica, replicas with the same seed identical, others not, and one replica equals the oracle"""
    R, n, w = 4096, 16384, 512.0
    ps = [abi.boids_params(radius=10.0, exact=0, seed=42 + (r % 2048)) for r in range(R)]
    b = kb.FlockerBatch((w, w), n, R, NORTH_STAR_DISC, True, ps, canonical_order=True)
    b.init()
    b.run(4)
    d = b.download()
    b.close()
    assert d["id"].shape == (R, n)
    assert (np.sort(d["id"], axis=1) == np.arange(n, dtype=d["id"].dtype)[None, :]).all()
    assert jump_or_still(d["ldx"], d["ldy"])
    for k in ("x", "y"):
        assert d[k].min() >= 0 and d[k].max() <= w
    for k in ("id", "x", "y", "ldx", "ldy"):
        assert (d[k][:2048].view(np.uint32) == d[k][2048:].view(np.uint32)).all(), k
    assert (d["x"][0] != d["x"][1]).any()
    r = 3000                                                        # seed 42 + 952
    op, _ = both_params(exact=0, seed=42 + (r % 2048))
    m = ob.Flockers(w, w, n, NORTH_STAR_DISC, True, op, canonical_order=True)
    m.init()
    m.step(4)
    x, y, dx, dy = m.agents()
    ids = d["id"][r]
    assert (d["x"][r].view(np.uint32) == x[ids].view(np.uint32)).all()
    assert (d["y"][r].view(np.uint32) == y[ids].view(np.uint32)).all()
    assert (d["ldx"][r].view(np.uint32) == dx[ids].view(np.uint32)).all()
    assert (d["ldy"][r].view(np.uint32) == dy[ids].view(np.uint32)).all()
