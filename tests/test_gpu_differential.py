"""Randomised differential tests (hypothesis) of the CUDA path against the oracle: arbitrary field
geometries, radii and populations — the generalisation of the reference's hand-written
known-answer tests.  Everything compared here is integer or set valued: bit-exact."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import krabmaga_b200 as kb
import oracle_binding as ob
from krabmaga_b200 import _abi as abi
from parity_util import bags, csr_lists

pytestmark = pytest.mark.gpu

# the gate runs a fixed example set; KG_FUZZ=1 explores fresh random examples on every run
FUZZ = os.environ.get("KG_FUZZ", "") not in ("", "0")
COMMON = dict(deadline=None, suppress_health_check=list(HealthCheck), derandomize=not FUZZ, database=None)

geometry = st.tuples(
    st.floats(3.0, 300.0, width=32), st.floats(3.0, 300.0, width=32),         # w, h
    st.sampled_from([0.5, 1.0, 2.5, 6.6666665, 10.0, 33.0]),                   # discretization
    st.booleans(),                                                              # toroidal
    st.integers(1, 400),                                                        # agents
    st.integers(0, 2**31 - 1))                                                  # seed


def population(w, h, n, seed):
    rng = np.random.default_rng(seed)
    x = (rng.random(n, dtype=np.float32) * np.float32(w)).astype(np.float32)
    y = (rng.random(n, dtype=np.float32) * np.float32(h)).astype(np.float32)
    # a few agents exactly on cell corners, on the far edges and on top of each other
    k = min(n, 6)
    x[:k] = np.array([0.0, w, w * 0.5, 0.0, w, x[-1]], np.float32)[:k]
    y[:k] = np.array([0.0, h, h * 0.5, h, 0.0, y[-1]], np.float32)[:k]
    return x, y


@settings(max_examples=200 if FUZZ else 40, **COMMON)
@given(geometry, st.floats(0.0, 60.0, width=32), st.booleans())
def test_rebuild_and_queries_match_the_oracle(geom, radius, exact):
    w, h, d, tor, n, seed = geom
    x, y = population(w, h, n, seed)
    ids = np.arange(n, dtype=np.uint32)
    z = np.zeros(n, np.float32)
    o = ob.Field2D(w, h, d, tor)
    try:
        o.set_object_locations(ids, x, y, z, z)
    except ob.OraclePanic:
        # the reference's Vec index panics: the ABI must refuse the same input
        f = kb.Field2D(w, h, d, tor, capacity=n)
        with pytest.raises(kb.KgOutOfBounds):
            f.set_object_locations(ids, x, y)
        f.close()
        return
    o.lazy_update()
    f = kb.Field2D(w, h, d, tor, capacity=n)
    f.set_order(True)
    f.set_object_locations(ids, x, y)
    f.lazy_update()
    assert (f.dw, f.dh) == o.dims()[:2]
    got, want = f.download(), o.iter_objects()
    assert bags(got) == bags(want)
    assert (f.cell_counts() == o.cell_counts()).all()
    rng = np.random.default_rng(seed ^ 0x5bd1e995)
    qx = (rng.random(12, dtype=np.float32) * np.float32(w)).astype(np.float32)
    qy = (rng.random(12, dtype=np.float32) * np.float32(h)).astype(np.float32)
    k = min(3, n)                                          # query at a few agents' own positions
    qx[:k], qy[:k] = x[:k], y[:k]
    offs, nb = f.neighbors_batch(np.stack([qx, qy], 1), float(radius), exact)
    woffs, wnb = o.neighbors_batch(qx, qy, float(radius), int(exact))
    assert csr_lists(offs, nb) == csr_lists(woffs, wnb)     # same ids in the same order
    f.close()


@settings(max_examples=60 if FUZZ else 15, **COMMON)
@given(st.integers(1, 70), st.integers(1, 9), st.integers(0, 2**31 - 1), st.integers(1, 12))
def test_forest_fire_random_grids(w, h16, seed, steps):
    """random states (trees, fire, ash, None) on ragged grid shapes, fast and generic kernel"""
    h = h16 * 16 if seed % 2 else h16 * 16 + seed % 15 + 1    # multiples of 16 take the fast kernel
    rng = np.random.default_rng(seed)
    cells = rng.choice(np.array([1, 1, 1, 2, 3, 0xFF], np.uint8), size=(w, h))
    o = ob.ForestFire(w, h)
    o.load(cells)
    o.step(steps)
    g = kb.DenseNumberGrid2D(w, h)
    g.upload(cells, unbuffered=True)
    g.lazy_update()
    g.run_stencil(steps)
    assert (g.download() == o.dump()).all()
    g.close()


step_case = st.tuples(
    st.floats(20.0, 260.0, width=32),                                          # w = h (bird.rs:146-147 wraps both axes by width)
    st.sampled_from([0.5, 1.0, 2.5, 6.6666665, 10.0, 33.0]),                   # discretization
    st.booleans(),                                                              # toroidal field
    st.integers(1, 500),                                                        # agents
    st.floats(0.5, 25.0, width=32),                                             # radius
    st.booleans(),                                                              # exact query
    st.integers(1, 4),                                                          # steps
    st.integers(0, 2**31 - 1))                                                  # seed
weights = st.tuples(*[st.floats(0.0, 3.0, width=32) for _ in range(5)], st.floats(0.0625, 1.5, width=32))


@settings(max_examples=150 if FUZZ else 30, **COMMON)
@given(step_case, weights)
def test_boids_steps_match_the_oracle_bit_for_bit(case, wts):
    """Bird::step for every agent (bird.rs:39-155) on arbitrary geometries, radii and weights,
    whichever K4 the dispatcher picks: with the same in-bag order on both sides every f32 of
    every agent must equal the oracle's after 1-4 steps."""
    w, d, tor, n, radius, exact, steps, seed = case
    coh, avo, rnd, con, mom, jump = (float(v) for v in wts)
    rng = np.random.default_rng(seed)
    x = (rng.random(n, dtype=np.float32) * np.float32(w)).astype(np.float32)
    y = (rng.random(n, dtype=np.float32) * np.float32(w)).astype(np.float32)
    # a crowd in one corner, twins on one spot, agents on the origin and just inside the far edge
    k = n // 3
    x[:k] = x[:k] * np.float32(0.05)
    y[:k] = y[:k] * np.float32(0.05)
    edge = np.nextafter(np.float32(w), np.float32(0))
    for i, (px, py) in enumerate([(0.0, 0.0), (edge, edge), (0.0, edge), (x[-1], y[-1])][:min(n, 4)]):
        x[i], y[i] = px, py
    ang = rng.random(n) * 2 * np.pi
    agents = dict(id=np.arange(n, dtype=np.uint32), x=np.minimum(x, edge), y=np.minimum(y, edge),
                  ldx=(0.7 * np.cos(ang)).astype(np.float32), ldy=(0.7 * np.sin(ang)).astype(np.float32))
    kw = dict(radius=float(radius), exact=int(exact), seed=seed, cohesion=coh, avoidance=avo, randomness=rnd,
              consistency=con, momentum=mom, jump=jump)
    m = ob.Flockers(w, w, n, d, tor, ob.boids_params(**kw), canonical_order=True)
    m.preset(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    m.init()
    m.step(steps)
    want = dict(zip(("x", "y", "ldx", "ldy"), m.agents()))
    state = kb.Flocker((w, w), n, discretization=d, toroidal=tor, params=abi.boids_params(**kw),
                       canonical_order=True, preset=agents)
    sch = kb.Schedule()
    state.init(sch)
    for _ in range(steps):
        sch.step_once(state)
    got = state.field1.download()
    ids = got["id"]
    assert (np.sort(ids) == agents["id"]).all()
    for key in ("x", "y", "ldx", "ldy"):
        bad = np.flatnonzero(got[key].view(np.uint32) != want[key][ids].view(np.uint32))
        assert len(bad) == 0, f"{key}: {len(bad)} of {n} differ, first ids {ids[bad[:5]]}"
    state.field1.close()


strip_case = st.tuples(
    st.sampled_from([6.6666665, 5.0, 10.0, 4.0]),                               # discretization
    st.integers(1, 3),                                                          # window half-width dd, in columns
    st.integers(2, 4),                                                          # strips
    st.integers(0, 60),                                                         # extra columns beyond the minimum
    st.floats(0.0, 0.875, width=32),                                            # fraction of a column on top
    st.integers(200, 3000),                                                     # agents
    st.integers(1, 12),                                                         # steps
    st.integers(0, 2**31 - 1))                                                  # seed


@settings(max_examples=80 if FUZZ else 20, **COMMON)
@given(strip_case, weights)
def test_strip_worlds_match_the_oracle_bit_for_bit(case, wts):
    """2-4 x-strips (halo exchange, ring migration, windows of 1-3 columns, any column width) ==
    the oracle's single world, every f32 of every agent, in the canonical in-bag order"""
    from krabmaga_b200 import strips
    d, dd, G, extra, frac, n, steps, seed = case
    coh, avo, rnd, con, mom, jump = (float(v) for v in wts)
    d32 = np.float32(d)
    radius = float(d32 * np.float32(dd + 0.5))
    cols = max(G * (dd + 2) + 2, 2 * (dd + 1) + 3) + extra
    w = float(d32 * np.float32(cols + frac))
    jump = min(jump, 0.9 * d)
    rng = np.random.default_rng(seed)
    edge = np.nextafter(np.float32(w), np.float32(0))
    x = np.minimum((rng.random(n, dtype=np.float32) * np.float32(w)).astype(np.float32), edge)
    y = np.minimum((rng.random(n, dtype=np.float32) * np.float32(w)).astype(np.float32), edge)
    # agents on every strip boundary, just left of it, and on the wrap-around seam
    bounds = [np.float32(x0) * d32 for x0, _ in strips.partition(w, w, d, G)]
    special = [b for b in bounds] + [np.nextafter(b, np.float32(0)) for b in bounds[1:]] + [edge]
    x[:len(special)] = np.array(special, np.float32)
    ang = rng.random(n) * 2 * np.pi
    agents = dict(id=np.arange(n, dtype=np.uint32), x=x, y=y,
                  ldx=(0.7 * np.cos(ang)).astype(np.float32), ldy=(0.7 * np.sin(ang)).astype(np.float32))
    kw = dict(radius=radius, exact=0, seed=seed, cohesion=coh, avoidance=avo, randomness=rnd,
              consistency=con, momentum=mom, jump=jump)
    m = ob.Flockers(w, w, n, d, True, ob.boids_params(**kw), canonical_order=True)
    m.preset(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    m.init()
    m.step(steps)
    want = dict(zip(("x", "y", "ldx", "ldy"), m.agents()))
    nd = kb._abi.lib().kg_device_count()
    world = strips.StripWorld(w, w, d, radius, [r % nd for r in range(G)], n, canonical_order=True, slack=4.0)
    world.upload(agents)
    gp = abi.boids_params(**kw)
    gp.step = 0
    world.run_boids(gp, steps)
    got = world.download()
    world.close()
    ids = got["id"]
    assert len(ids) == n and (np.sort(ids) == agents["id"]).all()
    for key in ("x", "y", "ldx", "ldy"):
        bad = np.flatnonzero(got[key].view(np.uint32) != want[key][ids].view(np.uint32))
        assert len(bad) == 0, f"{key}: {len(bad)} of {n} differ, first ids {ids[bad[:5]]}"


block_case = st.tuples(
    st.sampled_from([6.6666665, 5.0, 10.0, 4.0]),                               # discretization
    st.integers(1, 3),                                                          # window half-width dd, in cells
    st.integers(1, 4), st.integers(1, 3),                                       # blocks along x, along y
    st.integers(0, 40),                                                         # extra cells beyond the minimum
    st.floats(0.0, 0.875, width=32),                                            # fraction of a cell on top
    st.integers(200, 3000),                                                     # agents
    st.integers(1, 10),                                                         # steps
    st.booleans(),                                                              # exact query
    st.integers(0, 2**31 - 1))                                                  # seed


@settings(max_examples=80 if FUZZ else 20, **COMMON)
@given(block_case, weights)
def test_block_worlds_match_the_oracle_bit_for_bit(case, wts):
    """nbx x nby blocks (halo rings, corner neighbours, wrapping migrants, windows of 1-3 cells, any cell width) ==
    the oracle's single world, every f32 of every agent, in the canonical in-bag order"""
    from krabmaga_b200 import blocks
    d, dd, nbx, nby, extra, frac, n, steps, exact, seed = case
    coh, avo, rnd, con, mom, jump = (float(v) for v in wts)
    d32 = np.float32(d)
    radius = float(d32 * np.float32(dd + 0.5))
    cells = max(max(nbx, nby) * (dd + 2) + 2, 2 * (dd + 1) + 3) + extra
    w = float(d32 * np.float32(cells + frac))
    jump = min(jump, 0.9 * d)
    rng = np.random.default_rng(seed)
    edge = np.nextafter(np.float32(w), np.float32(0))
    x = np.minimum((rng.random(n, dtype=np.float32) * np.float32(w)).astype(np.float32), edge)
    y = np.minimum((rng.random(n, dtype=np.float32) * np.float32(w)).astype(np.float32), edge)
    # agents on block corners, just inside them, and on the wrap-around seams
    max_c = int(np.ceil(np.float32(w) / d32))
    bx = [np.float32(b * max_c // nbx) * d32 for b in range(nbx)]
    by = [np.float32(b * max_c // nby) * d32 for b in range(nby)]
    sx = [v for v in bx] + [np.nextafter(v, np.float32(0)) for v in bx[1:]] + [edge, np.float32(0)]
    sy = [v for v in by] + [np.nextafter(v, np.float32(0)) for v in by[1:]] + [edge, np.float32(0)]
    k = min(len(sx), len(sy), n)
    x[:k] = np.array(sx[:k], np.float32)
    y[:k] = np.array(sy[::-1][:k], np.float32)
    ang = rng.random(n) * 2 * np.pi
    agents = dict(id=np.arange(n, dtype=np.uint32), x=x, y=y,
                  ldx=(0.7 * np.cos(ang)).astype(np.float32), ldy=(0.7 * np.sin(ang)).astype(np.float32))
    kw = dict(radius=radius, exact=int(exact), seed=seed, cohesion=coh, avoidance=avo, randomness=rnd,
              consistency=con, momentum=mom, jump=jump)
    m = ob.Flockers(w, w, n, d, True, ob.boids_params(**kw), canonical_order=True)
    m.preset(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    m.init()
    m.step(steps)
    want = dict(zip(("x", "y", "ldx", "ldy"), m.agents()))
    nd = kb._abi.lib().kg_device_count()
    world = blocks.BlockWorld(w, w, d, radius, nbx, nby, [r % nd for r in range(nbx * nby)], n, canonical_order=True,
                              slack=6.0)
    world.upload(agents)
    gp = abi.boids_params(**kw)
    gp.step = 0
    world.run_boids(gp, steps)
    got = world.download()
    world.close()
    ids = got["id"]
    assert len(ids) == n and (np.sort(ids) == agents["id"]).all()
    for key in ("x", "y", "ldx", "ldy"):
        bad = np.flatnonzero(got[key].view(np.uint32) != want[key][ids].view(np.uint32))
        assert len(bad) == 0, f"{key}: {len(bad)} of {n} differ, first ids {ids[bad[:5]]}"


batch_case = st.tuples(
    st.floats(30.0, 200.0, width=32),                                           # w = h
    st.sampled_from([2.5, 6.6666665, 10.0]),                                    # discretization
    st.integers(1, 8),                                                          # replicas
    st.integers(1, 500),                                                        # agents per replica
    st.sampled_from(["relax", "exact", "mixed"]),                               # query kind over the batch
    st.booleans(),                                                              # one window size for all replicas?
    st.integers(1, 6),                                                          # steps
    st.integers(0, 2**31 - 1))                                                  # seed


@settings(max_examples=80 if FUZZ else 20, **COMMON)
@given(batch_case, st.lists(weights, min_size=8, max_size=8), st.lists(st.floats(0.0, 0.96875, width=32),
                                                                       min_size=8, max_size=8))
def test_batched_replicas_match_the_oracle_bit_for_bit(case, wts, fracs):
    """kg_batch_*: replicas with their own weights, radii and query kinds — one window size (packed
    kernels) or several (generic walk) — each equal to the oracle run of that replica alone"""
    w, d, R, n, kind, same_window, steps, seed = case
    ps = []
    for r in range(R):
        coh, avo, rnd, con, mom, jump = (float(v) for v in wts[r])
        cols = 1 if same_window else 1 + r % 3                  # window half-width in columns
        radius = float(np.float32(d) * np.float32(cols + fracs[r]))
        exact = {"relax": 0, "exact": 1, "mixed": r % 2}[kind]
        ps.append(dict(radius=radius, exact=exact, seed=(seed + 977 * r) % 2**31, cohesion=coh, avoidance=avo,
                       randomness=rnd, consistency=con, momentum=mom, jump=jump))
    b = kb.FlockerBatch((w, w), n, R, d, True, [abi.boids_params(**p) for p in ps], canonical_order=True)
    b.init()
    b.run(steps)
    got = b.download()
    b.close()
    for r in range(R):
        m = ob.Flockers(w, w, n, d, True, ob.boids_params(**ps[r]), canonical_order=True)
        m.init()
        m.step(steps)
        want = dict(zip(("x", "y", "ldx", "ldy"), m.agents()))
        ids = got["id"][r]
        assert (np.sort(ids) == np.arange(n)).all()
        for key in ("x", "y", "ldx", "ldy"):
            bad = np.flatnonzero(got[key][r].view(np.uint32) != want[key][ids].view(np.uint32))
            assert len(bad) == 0, f"replica {r} {key}: {len(bad)} of {n} differ"
