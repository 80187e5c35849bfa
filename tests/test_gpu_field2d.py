"""GPU parity tests for the Field2D path (rebuild, queries, fused boids step) through the C ABI,
checked against the oracle on the same seeded inputs.

Bars (BASELINE.json north_star): cell assignments and neighbour sets bit-exact; per-step f32
positions within 1e-5 relative (summation order may differ); with KG_ORDER_CANONICAL on both
sides the summation order is identical and positions are required to be bit-exact.
"""
import numpy as np
import pytest

import krabmaga_b200 as kb
import oracle_binding as ob
from krabmaga_b200 import _abi as abi
from parity_util import (NORTH_STAR_DISC, bags, both_params, by_id, cells_by_id, csr_lists, csr_sets,
                         random_agents, rel_err)

pytestmark = pytest.mark.gpu

WIDTH, HEIGHT, DISC, TOROIDAL = 10.0, 10.0, 0.5, True

GEOMS = [  # (w, h, disc, toroidal, n)
    (10.0, 10.0, 0.5, True, 300),              # the fixture (dd = 20 for radius 10)
    (400.0, 400.0, NORTH_STAR_DISC, True, 10000),   # BASELINE config 1 (dd = 1)
    (400.0, 400.0, NORTH_STAR_DISC, False, 4000),   # non-toroidal: window wraps via t_transform
    (123.0, 77.0, 3.7, True, 5000),            # non-square, ragged cell sizes
    (64.0, 64.0, 8.0, True, 2000),             # w/disc integral: no partial last cell
]


def make_pair(w, h, d, t, agents, canonical=False):
    o = ob.Field2D(w, h, d, t)
    o.set_object_locations(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    o.lazy_update()
    g = kb.Field2D(w, h, d, t, capacity=len(agents["id"]) + 8)
    g.set_order(canonical)
    g.set_object_locations(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    g.lazy_update()
    return o, g


# ---------------------------------------------------------------- reference known-answer tests
def test_field_2d_neighbors_kat():
    """tests/engine/field_2d.rs:58-117 on the GPU field"""
    for fly in (5.0, 6.0, 7.0, 8.0, 9.0):
        f = kb.Field2D(WIDTH, HEIGHT, DISC, TOROIDAL, capacity=16)
        f.set_object_location((1, 0, 0), (0.0, 0.0))
        f.set_object_location((2, 0, 0), (0.0, 0.0))
        f.lazy_update()
        assert f.nagents == 2
        assert len(f.get_neighbors_within_distance((5.0, 5.0), 1.0)) == 0
        assert len(f.get_neighbors_within_relax_distance((5.0, 5.0), 1.0)) == 0
        f.set_object_location((1, 0, 0), (fly, fly))
        f.set_object_location((2, 0, 0), (0.0, 0.0))
        f.lazy_update()
        assert list(f.get_neighbors_within_distance((fly, fly), 1.0)) == [1]
        assert list(f.get_neighbors_within_distance((0.0, 0.0), 1.0)) == [2]
        assert sorted(f.get_neighbors_within_distance((5.0, 5.0), 10.0)) == [1, 2]
        assert sorted(f.get_neighbors_within_relax_distance((5.0, 5.0), 10.0)) == [1, 2]
        f.close()


def test_field_2d_gets_kat():
    """tests/engine/field_2d.rs:125-178"""
    f = kb.Field2D(WIDTH, HEIGHT, DISC, TOROIDAL, capacity=16)
    f.set_object_location((1,), (0.0, 0.0))
    f.set_object_location((2,), (5.0, 5.0))
    f.set_object_location((3,), (5.0, 5.0))
    f.lazy_update()
    assert f.nagents == 3
    assert sorted(f.get_objects((5.0, 5.0))) == [2, 3]
    assert len(f.get_objects((10.0, 0.0))) == 0  # padding column exists (:150-151)
    assert f.num_objects_at_location((5.0, 5.0)) == 2
    assert f.num_objects_at_location((0.0, 0.0)) == 1
    f.set_object_location((4,), (0.0, 0.0))
    assert len(f.get_objects_unbuffered((0.0, 0.0))) == 1
    assert len(f.get_objects((0.0, 0.0))) == 1
    f.remove_object_location((4,), (0.0, 0.0))
    assert len(f.get_objects_unbuffered((0.0, 0.0))) == 0
    f.lazy_update()
    assert f.num_objects() == 0


def test_field_2d_bags_kat():
    """tests/engine/field_2d.rs:186-210"""
    f = kb.Field2D(10.0, 10.0, DISC, TOROIDAL, capacity=16)
    assert (f.dw, f.dh) == (21, 21)
    assert f.num_empty_bags() == f.dh * f.dw == 441
    f.set_object_location((1,), (0.0, 0.0))
    f.set_object_location((2,), (0.0, 0.0))
    f.set_object_location((3,), (4.0, 4.0))
    f.lazy_update()
    assert f.num_empty_bags() == f.dh * f.dw - 2
    empty = f.get_empty_bags()
    assert len(empty) == 439 and (0.0, 0.0) not in empty and (4.0, 4.0) not in empty
    # get_random_empty_bag (:764-772, the doctest's property): always an empty bag's origin
    import random
    rng = random.Random(7)
    picks = {tuple(f.get_random_empty_bag(rng)) for _ in range(5)}
    assert picks <= {tuple(e) for e in empty} and len(picks) > 1


def test_field_2d_iter_kat():
    """tests/engine/field_2d.rs:218-247"""
    f = kb.Field2D(10.0, 10.0, DISC, TOROIDAL, capacity=16)
    f.set_object_location((1,), (0.0, 0.0))
    f.set_object_location((2,), (0.01, 0.01))
    f.set_object_location((3,), (5.0, 5.0))
    seen = []
    f.iter_objects(lambda loc, oid: seen.append(oid in f.get_objects_unbuffered(loc)), unbuffered=True)
    assert seen == [True] * 3
    f.lazy_update()
    assert len(f.get_objects((0.0, 0.0))) == 2
    seen = []
    f.iter_objects(lambda loc, oid: seen.append(oid in f.get_objects(loc)))
    assert seen == [True] * 3


def test_field_2d_single_step_kat():
    """tests/engine/field_2d.rs:31-50 through Flocker + Schedule on the GPU field"""
    state = kb.Flocker((WIDTH, HEIGHT), 10)
    schedule = kb.Schedule()
    state.init(schedule)
    schedule.step_once(state)
    assert state.field1.nagents == 10
    assert len(state.field1.get_neighbors_within_distance((5.0, 5.0), 10.0)) == 10
    assert len(state.field1.get_neighbors_within_relax_distance((5.0, 5.0), 10.0)) == 10


def test_out_of_world_is_reported_not_ignored():
    """field_2d.rs:838-842 index panic -> KG_E_OOB, nothing appended"""
    f = kb.Field2D(10.0, 10.0, DISC, TOROIDAL, capacity=16)
    with pytest.raises(kb.KgOutOfBounds):
        f.set_object_location((1,), (-1.0, 0.0))
    with pytest.raises(kb.KgOutOfBounds):
        f.set_object_location((1,), (100.0, 100.0))
    assert f.num_objects(unbuffered=True) == 0
    f.set_object_location((1,), (1.0, 1.0))
    f.lazy_update()
    assert f.num_objects() == 1


def test_capacity_is_enforced():
    f = kb.Field2D(10.0, 10.0, DISC, TOROIDAL, capacity=4)
    with pytest.raises(kb.KgError) as e:
        f.set_object_locations(np.arange(5), np.ones(5), np.ones(5))
    assert e.value.code == abi.KG_E_CAPACITY


def test_empty_field():
    f = kb.Field2D(10.0, 10.0, DISC, TOROIDAL, capacity=4)
    f.lazy_update()
    assert f.num_objects() == 0 and f.num_empty_bags() == 441
    assert len(f.get_neighbors_within_relax_distance((5.0, 5.0), 10.0)) == 0
    f.step_boids(abi.boids_params())
    f.lazy_update()
    assert f.num_objects() == 0


# ---------------------------------------------------------------- rebuild parity (K1-K3)
@pytest.mark.parametrize("w,h,d,t,n", GEOMS)
def test_rebuild_cells_and_bags_bit_exact(w, h, d, t, n):
    agents = random_agents(n, w, h, seed=n)
    o, g = make_pair(w, h, d, t, agents)
    dw, dh, _, _ = o.dims()
    assert (g.dw, g.dh) == (dw, dh)
    od, gd = o.iter_objects(), g.download()
    assert len(gd["id"]) == n
    assert (cells_by_id(od) == cells_by_id(gd)).all()          # cell assignment, per agent
    assert (np.diff(gd["cell"]) >= 0).all()                     # iter_objects order = cell-major
    assert bags(od) == bags(gd)                                 # bag membership
    assert (o.cell_counts()[: dw * dh] == g.cell_counts()).all()
    go = by_id(gd)
    for k in ("x", "y", "ldx", "ldy"):                          # payload carried bit for bit
        assert (go[k] == agents[k]).all()


def test_edge_coordinates_land_in_padding_cells():
    """x == w / y == h are legal (toroidal_transform can return dim) and use the +1 row/column"""
    w = h = 400.0
    d = NORTH_STAR_DISC
    agents = dict(id=np.arange(6, dtype=np.uint32),
                  x=np.array([400.0, 0.0, 400.0, 399.99997, 6.6666665, 13.333333], np.float32),
                  y=np.array([0.0, 400.0, 400.0, 399.99997, 6.6666665, 13.333333], np.float32),
                  ldx=np.zeros(6, np.float32), ldy=np.zeros(6, np.float32))
    o, g = make_pair(w, h, d, True, agents)
    assert (cells_by_id(o.iter_objects()) == cells_by_id(g.download())).all()
    offs, ids = g.neighbors_batch(np.stack([agents["x"], agents["y"]], 1), 10.0, exact=False)
    ooffs, oids = o.neighbors_batch(agents["x"], agents["y"], 10.0, 0)
    assert csr_sets(offs, ids) == csr_sets(ooffs, oids)


def test_canonical_order_sorts_every_bag_by_id():
    agents = random_agents(5000, 50.0, 50.0, seed=3)  # dense: ~20 per cell at disc 3
    agents["id"] = np.random.default_rng(5).permutation(5000).astype(np.uint32)
    _, g = make_pair(50.0, 50.0, 3.0, True, agents, canonical=True)
    d = g.download()
    same = np.diff(d["cell"]) == 0
    assert (np.diff(d["id"].astype(np.int64))[same] > 0).all()


# ---------------------------------------------------------------- neighbour-set parity (rows E/F)
@pytest.mark.parametrize("w,h,d,t,n", GEOMS)
@pytest.mark.parametrize("exact", [False, True])
def test_neighbor_sets_bit_exact(w, h, d, t, n, exact):
    agents = random_agents(n, w, h, seed=7 * n + exact)
    o, g = make_pair(w, h, d, t, agents, canonical=True)
    rng = np.random.default_rng(11)
    nq = 400
    qx = np.concatenate([agents["x"][:nq // 2], (rng.random(nq // 2) * w).astype(np.float32)])
    qy = np.concatenate([agents["y"][:nq // 2], (rng.random(nq // 2) * h).astype(np.float32)])
    for dist in (10.0, 2.5, d, 0.0, -1.0):
        offs, ids = g.neighbors_batch(np.stack([qx, qy], 1), float(dist), exact)
        ooffs, oids = o.neighbors_batch(qx, qy, float(dist), int(exact))
        assert (offs == ooffs).all(), f"counts differ at dist={dist}"
        assert csr_sets(offs, ids) == csr_sets(ooffs, oids)
        # the window walk order (x asc, y asc) is the reference's: compare the cell sequence
        cell_of = cells_by_id(g.download())
        for a, b in zip(csr_lists(offs, ids)[:50], csr_lists(ooffs, oids)[:50]):
            assert [cell_of[i] for i in a] == [cell_of[i] for i in b]


def test_neighbor_query_capacity_retry():
    agents = random_agents(2000, 10.0, 10.0, seed=1)
    _, g = make_pair(10.0, 10.0, 0.5, True, agents)
    offs, ids = g.neighbors_batch(np.full((40, 2), 5.0, np.float32), 10.0, exact=False)
    assert offs[-1] == len(ids) == 40 * 2000  # every query sees every agent


# ---------------------------------------------------------------- fused step parity (row G)
def run_oracle_steps(w, h, d, t, agents, oparams, nsteps, canonical):
    m = ob.Flockers(w, h, len(agents["id"]), d, t, oparams, canonical_order=canonical)
    m.preset(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    m.init()
    m.step(nsteps)
    x, y, dx, dy = m.agents()
    return dict(x=x, y=y, ldx=dx, ldy=dy), m


def run_gpu_steps(w, h, d, t, agents, gparams, nsteps, canonical):
    st = kb.Flocker((w, h), len(agents["id"]), discretization=d, toroidal=t, params=gparams,
                    canonical_order=canonical, preset=agents)
    sch = kb.Schedule()
    st.init(sch)
    for _ in range(nsteps):
        sch.step_once(st)
    return by_id(st.field1.download()), st


STEP_CASES = [  # (w, h, disc, toroidal, n, exact)
    (400.0, 400.0, NORTH_STAR_DISC, True, 10000, 0),   # config 1, north-star geometry, relax
    (400.0, 400.0, NORTH_STAR_DISC, True, 10000, 1),   # same, exact query
    (10.0, 10.0, 0.5, True, 64, 1),                    # fixture geometry: 41x41 window, exact
    (10.0, 10.0, 0.5, True, 64, 0),
    (200.0, 200.0, 0.5, True, 100, 1),                 # tests/explore/simulate.rs geometry
    (400.0, 400.0, NORTH_STAR_DISC, False, 3000, 0),   # non-toroidal field
    (90.0, 90.0, 4.5, True, 3000, 1),                  # dd = 2
]


@pytest.mark.parametrize("w,h,d,t,n,exact", STEP_CASES)
def test_one_step_positions_within_1e5_any_order(w, h, d, t, n, exact):
    """default (atomic) bag order vs the reference's schedule order: same sets, different
    summation order -> <= 1e-5 relative to the world size"""
    agents = random_agents(n, w, h, seed=100 + n + exact)
    op, gp = both_params(radius=10.0, exact=exact, seed=42)
    want, _ = run_oracle_steps(w, h, d, t, agents, op, 1, canonical=False)
    got, _ = run_gpu_steps(w, h, d, t, agents, gp, 1, canonical=False)
    # tolerance from BASELINE.json north_star: 1e-5 relative (positions compared on the torus)
    for k, dim in (("x", w), ("y", w)):
        diff = np.abs(got[k].astype(np.float64) - want[k].astype(np.float64))
        diff = np.minimum(diff, dim - diff)
        assert (diff <= 1e-5 * np.maximum(np.abs(want[k]), 1.0)).all(), float(diff.max())
    # last_d = JUMP * d/|d|: when the five force terms nearly cancel, |d| is small and the
    # normalisation amplifies summation-order noise, so a handful of agents may exceed 1e-5 here
    # while their positions (the north-star criterion, above) stay far inside it.  Require 99.9 %
    # within 1e-5 and every agent within 1e-4.
    for k in ("ldx", "ldy"):
        diff = np.abs(got[k].astype(np.float64) - want[k].astype(np.float64))
        assert (diff <= 1e-4).all(), float(diff.max())
        assert (diff <= 1e-5).mean() >= 0.999, float((diff > 1e-5).mean())


@pytest.mark.parametrize("w,h,d,t,n,exact", STEP_CASES)
def test_steps_bit_exact_in_canonical_order(w, h, d, t, n, exact):
    """identical summation order on both sides -> every f32 must match bit for bit, and keep
    matching as the trajectories evolve"""
    agents = random_agents(n, w, h, seed=200 + n + exact)
    op, gp = both_params(radius=10.0, exact=exact, seed=7)
    nsteps = 20 if n >= 10000 else 40
    want, _ = run_oracle_steps(w, h, d, t, agents, op, nsteps, canonical=True)
    got, _ = run_gpu_steps(w, h, d, t, agents, gp, nsteps, canonical=True)
    for k in ("x", "y", "ldx", "ldy"):
        bad = np.flatnonzero(got[k].view(np.uint32) != want[k].view(np.uint32))
        assert len(bad) == 0, f"{k}: {len(bad)} of {n} differ, first id {bad[:5]}"


def test_config1_200_steps_bit_exact_canonical():
    """BASELINE config 1: 10,000 agents, 400x400 toroidal, radius 10, 200 steps, fixed seed,
    Philox init on both sides"""
    w = h = 400.0
    n = 10000
    op, gp = both_params(radius=10.0, exact=0, seed=42)
    m = ob.Flockers(w, h, n, NORTH_STAR_DISC, True, op, canonical_order=True)
    m.init()
    m.step(200)
    x, y, dx, dy = m.agents()
    st = kb.Flocker((w, h), n, discretization=NORTH_STAR_DISC, params=gp, canonical_order=True)
    kb.simulate(st, 200, 1)
    got = by_id(st.field1.download())
    assert (got["x"].view(np.uint32) == x.view(np.uint32)).all()
    assert (got["y"].view(np.uint32) == y.view(np.uint32)).all()
    assert (got["ldx"].view(np.uint32) == dx.view(np.uint32)).all()
    assert (got["ldy"].view(np.uint32) == dy.view(np.uint32)).all()


def test_philox_init_bit_exact():
    """State::init (state.rs:41-56) with the shared Philox stream"""
    op, gp = both_params(seed=1234567890123)
    m = ob.Flockers(400.0, 400.0, 5000, NORTH_STAR_DISC, True, op)
    m.init()
    od = m.field1.iter_objects(unbuffered=True)
    f = kb.Field2D(400.0, 400.0, NORTH_STAR_DISC, True, capacity=5000)
    f.init_flockers(5000, 1234567890123)
    gd = f.download(unbuffered=True)
    o, g = by_id(od), by_id(gd)
    for k in ("x", "y", "ldx", "ldy"):
        assert (o[k].view(np.uint32) == g[k].view(np.uint32)).all()


def test_run_boids_equals_stepwise_calls():
    agents = random_agents(4000, 400.0, 400.0, seed=9)
    _, gp = both_params(exact=0, seed=3)
    a, _ = run_gpu_steps(400.0, 400.0, NORTH_STAR_DISC, True, agents, gp, 7, canonical=True)
    f = kb.Field2D(400.0, 400.0, NORTH_STAR_DISC, True, capacity=4000)
    f.set_order(True)
    f.set_object_locations(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    f.lazy_update()
    gp.step = 0
    f.run_boids(gp, 7)
    b = by_id(f.download())
    for k in a:
        assert (a[k].view(np.uint32) == b[k].view(np.uint32)).all()


def test_step_boids_host_roundtrip():
    """the e2e entry point: host SoA in -> host SoA out equals upload/step/download"""
    n = 3000
    agents = random_agents(n, 400.0, 400.0, seed=21)
    _, gp = both_params(exact=0, seed=5)
    f = kb.Field2D(400.0, 400.0, NORTH_STAR_DISC, True, capacity=n)
    f.set_order(True)
    out = {k: np.zeros(n, v.dtype) for k, v in agents.items()}
    f.step_boids_host(gp, agents, out)
    want, _ = run_gpu_steps(400.0, 400.0, NORTH_STAR_DISC, True, agents, gp, 1, canonical=True)
    got = by_id(out)
    for k in want:
        assert (got[k].view(np.uint32) == want[k].view(np.uint32)).all()


@pytest.mark.parametrize("implicit_ids", [False, True])
def test_step_boids_host_ordered_equals_the_keyed_entry(implicit_ids):
    """the by-position e2e entry: input agent i comes back at index i, bit for bit what the keyed
    entry (ids travel both ways) returns for that id; with and without an id array; canonical order
    so that both runs sum their neighbours alike"""
    n = 5000
    agents = random_agents(n, 400.0, 400.0, seed=33)
    if implicit_ids:
        agents["id"] = np.arange(n, dtype=np.uint32)
    else:
        agents["id"] = np.random.default_rng(5).permutation(n).astype(np.uint32) * 3 + 7   # sparse, shuffled
    _, gp = both_params(exact=0, seed=9)
    f = kb.Field2D(400.0, 400.0, NORTH_STAR_DISC, True, capacity=n)
    f.set_order(True)
    keyed = {k: np.zeros(n, v.dtype) for k, v in agents.items()}
    f.step_boids_host(gp, agents, keyed)
    where = {int(i): j for j, i in enumerate(keyed["id"])}
    sel = np.array([where[int(i)] for i in agents["id"]])
    inp = {k: agents[k].copy() for k in ("x", "y", "ldx", "ldy")}
    if not implicit_ids:
        inp["id"] = agents["id"]
    out = {k: np.zeros(n, np.float32) for k in ("x", "y", "ldx", "ldy")}
    f.step_boids_host_ordered(gp, inp, out)
    for k in out:
        assert (out[k].view(np.uint32) == keyed[k][sel].view(np.uint32)).all(), k
    # the four arrays as slices of one block on both sides (the entry then moves one copy each way)
    blk_in, blk_out = np.empty(4 * n, np.float32), np.zeros(4 * n, np.float32)
    vin = {k: blk_in[j * n:(j + 1) * n] for j, k in enumerate(("x", "y", "ldx", "ldy"))}
    vout = {k: blk_out[j * n:(j + 1) * n] for j, k in enumerate(("x", "y", "ldx", "ldy"))}
    for k in vin:
        vin[k][:] = agents[k]
    if not implicit_ids:
        vin["id"] = agents["id"]
    f.step_boids_host_ordered(gp, vin, vout)
    for k in out:
        assert (vout[k].view(np.uint32) == out[k].view(np.uint32)).all(), k
    # in place (out aliases in), twice: step 2 continues from step 1's result
    gp2 = kb.boids_params(radius=10.0, exact=0, seed=9)
    gp2.step = gp.step + 1
    f.step_boids_host_ordered(gp, inp, inp)
    for k in out:
        assert (inp[k].view(np.uint32) == out[k].view(np.uint32)).all(), k
    f.step_boids_host_ordered(gp2, inp, inp)
    nxt = dict(id=agents["id"], **out)
    keyed2 = {k: np.zeros(n, v.dtype) for k, v in agents.items()}
    f.step_boids_host(gp2, nxt, keyed2)
    where = {int(i): j for j, i in enumerate(keyed2["id"])}
    sel = np.array([where[int(i)] for i in agents["id"]])
    for k in out:
        assert (inp[k].view(np.uint32) == keyed2[k][sel].view(np.uint32)).all(), k


def test_step_boids_host_ordered_any_order_and_errors():
    """default bag order: same agents within the north-star's 1e-5; out-of-grid input is refused"""
    n = 4000
    agents = random_agents(n, 400.0, 400.0, seed=34)
    _, gp = both_params(exact=0, seed=3)
    f = kb.Field2D(400.0, 400.0, NORTH_STAR_DISC, True, capacity=n)
    inp = {k: agents[k].copy() for k in ("x", "y", "ldx", "ldy")}
    out = {k: np.zeros(n, np.float32) for k in inp}
    f.step_boids_host_ordered(gp, inp, out)
    g = kb.Field2D(400.0, 400.0, NORTH_STAR_DISC, True, capacity=n)
    g.set_order(True)
    ref = {k: np.zeros(n, np.float32) for k in inp}
    g.step_boids_host_ordered(gp, inp, ref)
    for k, dim in (("x", 400.0), ("y", 400.0)):
        diff = np.abs(out[k].astype(np.float64) - ref[k].astype(np.float64))
        diff = np.minimum(diff, dim - diff)
        assert (diff <= 1e-5 * np.maximum(np.abs(ref[k]), 1.0)).all()
    bad = {k: v.copy() for k, v in inp.items()}
    bad["x"][17] = 1.0e6
    with pytest.raises(Exception):
        f.step_boids_host_ordered(gp, bad, out)


# ---------------------------------------------------------------- full-size properties
def test_1m_agents_size_independent_properties():
    """BASELINE config 2 (1M agents, 4000x4000): what can be checked without the oracle"""
    n, w = 1_000_000, 4000.0
    _, gp = both_params(exact=0, seed=42)
    f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=n)
    f.init_flockers(n, 42)
    f.lazy_update()
    before = f.download()
    f.run_boids(gp, 5)
    d = f.download()
    assert len(d["id"]) == n
    assert (np.sort(d["id"]) == np.arange(n)).all()              # agents conserved, no duplicates
    assert (np.diff(d["cell"]) >= 0).all()                        # sorted by cell
    counts = f.cell_counts()
    assert counts.sum() == n
    assert (np.bincount(d["cell"], minlength=len(counts)) == counts).all()
    assert (d["x"] >= 0).all() and (d["x"] <= w).all() and (d["y"] >= 0).all() and (d["y"] <= w).all()
    norm = np.hypot(d["ldx"].astype(np.float64), d["ldy"].astype(np.float64))
    assert np.allclose(norm[norm > 0], 0.7, rtol=1e-5)            # |last_d| == JUMP
    # each step moves an agent by exactly JUMP (toroidally): 5 steps -> at most 3.5
    b, a = by_id(before), by_id(d)
    dx = np.abs(a["x"] - b["x"]); dx = np.minimum(dx, w - dx)
    dy = np.abs(a["y"] - b["y"]); dy = np.minimum(dy, w - dy)
    assert np.hypot(dx, dy).max() <= 3.5 + 1e-3
    # spot-check 2000 agents' neighbour sets against the oracle at full size
    o = ob.Field2D(w, w, NORTH_STAR_DISC, True)
    o.set_object_locations(d["id"], d["x"], d["y"], d["ldx"], d["ldy"])
    o.lazy_update()
    q = np.random.default_rng(0).choice(n, 2000, replace=False)
    offs, ids = f.neighbors_batch(np.stack([d["x"][q], d["y"][q]], 1), 10.0, exact=False)
    ooffs, oids = o.neighbors_batch(d["x"][q], d["y"][q], 10.0, 0)
    assert csr_sets(offs, ids) == csr_sets(ooffs, oids)


# ---------------------------------------------------------------- fast K4 vs generic K4
def test_shared_reciprocal_division_is_ieee_exact():
    """fdiv2_shared (two quotients from one refined reciprocal) == __fdiv_rn, 2^28 random triples
    over the operand domain the fast kernel can produce"""
    import ctypes as C
    bad = C.c_uint64(123)
    abi.check(abi.lib().kg_selftest_div(0, 1 << 28, 2024, C.byref(bad)))
    assert bad.value == 0


EXACT_VARIANTS = (abi.KG_K4_GENERIC, abi.KG_K4_AUTO, abi.KG_K4_PACKED_BY_ID)


def _exact_case(radius):
    n, w = 12000, 420.0
    agents = random_agents(n, w, w, seed=int(radius * 10))
    agents["x"][:6] = [0.0, 1e-7, 5e-7, w - 1e-4, 6.6666665, 13.333333]   # origin, edges, cell corners
    agents["y"][:6] = [1e-8, 0.0, 3.0, 2e-7, 6.6666665, 13.333333]
    _, gp = both_params(radius=radius, exact=1, seed=19, cohesion=1.3, avoidance=0.9, consistency=0.7,
                        randomness=1.7, momentum=1.1, jump=0.65)
    return n, w, agents, gp


@pytest.mark.parametrize("radius", [10.0, 6.0, 13.3, 0.5, 14.0, 21.0, 33.0])
def test_packed_exact_query_kernel_equals_the_generic_kernel(radius):
    """get_neighbors_within_distance on the packed path (cell classes by corner test, per-element
    threshold on the bit pattern of dx^2+dy^2) vs the reference-shaped generic walk, bit for bit over
    25 steps: a 3x3 window (radius 10, 13.3), one whose corner cells are always skipped (6.0), a
    one-cell window (0.5) and 5x5 / 7x7 / 9x9 windows (14, 21, 33: columns walked in groups of three
    cells); non-unit weights, edge and near-origin agents"""
    n, w, agents, gp = _exact_case(radius)
    outs = {}
    for variant in EXACT_VARIANTS:
        f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=n)
        f.set_order(True)
        f.set_kernel_variant(variant)
        f.set_object_locations(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
        f.lazy_update()
        gp.step = 0
        f.run_boids(gp, 25)
        outs[variant] = by_id(f.download())
        f.close()
    for variant, out in outs.items():
        for k in out:
            assert (out[k].view(np.uint32) == outs[abi.KG_K4_GENERIC][k].view(np.uint32)).all(), (variant, k)


def test_packed_exact_query_with_duplicate_ids():
    """duplicate ids: every variant steps the SAME read buffer once (bag order is then irrelevant)"""
    n, w, agents, gp = _exact_case(10.0)
    ids = agents["id"].copy()
    ids[100:200] = ids[300:400]
    f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=4 * n)
    f.set_object_locations(ids, agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    f.lazy_update()
    outs = []
    for variant in EXACT_VARIANTS:
        f.set_kernel_variant(variant)
        gp.step = 0
        f.step_boids(gp)
        d = f.download(unbuffered=True, with_cells=False)
        outs.append({k: v[-n:].copy() for k, v in d.items()})
    f.close()
    for o in outs[1:]:
        for k in o:
            assert (o[k].view(np.uint32) == outs[0][k].view(np.uint32)).all(), k


@pytest.mark.parametrize("n,w", [(10000, 400.0), (60000, 900.0)])
def test_every_k4_variant_equals_the_generic_kernel(n, w):
    """the specialised kernels (toroidal, relax, 3x3 window: packed FADD2/FFMA2 loop with index or
    id self-exclusion, and the scalar fast kernel) must reproduce the generic window walk bit for
    bit over many steps"""
    agents = random_agents(n, w, w, seed=n)
    agents["x"][:4] = [0.0, 1e-7, 5e-7, w - 1e-4]      # exercise the near-origin (unsafe) guard
    agents["y"][:4] = [1e-8, 0.0, 3.0, 2e-7]
    # weights that are not 1.0: a multiply-add contraction anywhere in a kernel (x*1+y == x+y would
    # hide it) changes bits here
    _, gp = both_params(exact=0, seed=11, cohesion=1.3, avoidance=0.9, consistency=0.7,
                        randomness=1.7, momentum=1.1, jump=0.65)
    outs = {}
    for variant in (abi.KG_K4_GENERIC, abi.KG_K4_AUTO, abi.KG_K4_FAST_SCALAR, abi.KG_K4_PACKED_BY_ID,
                    abi.KG_K4_TILED, abi.KG_K4_STAGED):
        f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=n)
        f.set_order(True)
        f.set_kernel_variant(variant)
        f.set_object_locations(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
        f.lazy_update()
        gp.step = 0
        f.run_boids(gp, 30)
        outs[variant] = by_id(f.download())
        f.close()
    for variant, out in outs.items():
        for k in out:
            assert (out[k].view(np.uint32) == outs[abi.KG_K4_GENERIC][k].view(np.uint32)).all(), \
                (variant, k)


def test_any_order_one_step_all_variants_agree_on_the_same_read_buffer():
    """KG_ORDER_ANY: whatever order the scatter left the bags in, every K4 variant must produce
    the same bits from that same read buffer (they walk it in the same order)"""
    n, w = 20000, 500.0
    agents = random_agents(n, w, w, seed=5)
    _, gp = both_params(exact=0, seed=3)
    f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=6 * n)   # six steps into one write log
    f.set_object_locations(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    f.lazy_update()
    f.run_boids(gp, 3)
    outs = []
    for variant in (abi.KG_K4_GENERIC, abi.KG_K4_AUTO, abi.KG_K4_FAST_SCALAR, abi.KG_K4_PACKED_BY_ID,
                    abi.KG_K4_TILED, abi.KG_K4_STAGED):
        f.set_kernel_variant(variant)
        gp.step = 3
        f.step_boids(gp)                       # appends to the write log; the read buffer stays
        d = f.download(unbuffered=True, with_cells=False)
        outs.append({k: v[-n:].copy() for k, v in d.items()})
    for o in outs[1:]:
        for k in o:
            assert (o[k].view(np.uint32) == outs[0][k].view(np.uint32)).all(), k
    f.close()


def test_duplicate_ids_fall_back_to_the_id_comparison():
    """bird.rs:63 skips every neighbour whose id equals the agent's own.  The packed kernel may
    only use its index test after verifying the ids unique; with duplicates it must compare ids
    and match the oracle bit for bit."""
    n, w = 4000, 250.0
    agents = random_agents(n, w, w, seed=77)
    agents["id"][:] = np.arange(n, dtype=np.uint32) // 2          # every id appears twice
    # ... on two agents in neighbouring cells (different bags, so that the canonical in-bag order
    # by id has no ties), close enough to be in each other's 3x3 window
    agents["x"][0::2] *= np.float32((w - 8.0) / w)
    agents["x"][1::2] = agents["x"][0::2] + np.float32(7.0)
    agents["y"][1::2] = agents["y"][0::2]
    op, gp = both_params(exact=0, seed=5)
    m = ob.Flockers(w, w, n, NORTH_STAR_DISC, True, op, canonical_order=True)
    m.preset(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    m.init()
    m.step(1)
    want = m.field1.iter_objects()
    f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=n)
    f.set_order(True)
    f.set_object_locations(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    f.lazy_update()
    gp.step = 0
    f.run_boids(gp, 1)
    got = f.download()
    f.close()
    # twins share an id: compare as multisets of (id, x, y, ldx, ldy) bit patterns
    def rows(d):
        r = np.stack([d["id"].astype(np.uint32), d["x"].view(np.uint32), d["y"].view(np.uint32),
                      d["ldx"].view(np.uint32), d["ldy"].view(np.uint32)], axis=1)
        return r[np.lexsort(r.T[::-1])]
    assert (rows(got) == rows(want)).all()


def test_ids_beyond_the_bitmap_still_step_correctly():
    """ids too large to verify unique select the id-comparing loop; results must not change.
    (The Philox stream is keyed by id, so the randomness weight is switched off here.)"""
    n, w = 3000, 200.0
    agents = random_agents(n, w, w, seed=9)
    gp = abi.boids_params(exact=0, seed=2, randomness=0.0)
    outs = []
    for offset in (0, 0xF0000000):
        f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=n)
        f.set_order(True)
        f.set_object_locations(agents["id"] + np.uint32(offset), agents["x"], agents["y"],
                               agents["ldx"], agents["ldy"])
        f.lazy_update()
        gp.step = 0
        f.run_boids(gp, 5)
        d = f.download()
        o = np.argsort(d["id"])
        outs.append({k: d[k][o] for k in ("x", "y", "ldx", "ldy")})
        f.close()
    for k in outs[0]:
        assert (outs[0][k].view(np.uint32) == outs[1][k].view(np.uint32)).all(), k


def test_queries_return_the_neighbours_themselves_and_reuse_their_scratch():
    """field_2d.rs:386 / :472 return Vec<O>: kg_field2d_neighbors_agents hands back (pos, last_d) of
    every neighbour beside its id; repeated calls (the per-agent pattern of a model that keeps a CPU
    step) must agree with the id-only entry point and with the uploaded state"""
    n, w = 5000, 300.0
    agents = random_agents(n, w, w, seed=77)
    f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=n)
    f.set_order(True)
    f.set_object_locations(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    f.lazy_update()
    locs = np.stack([agents["x"][:300], agents["y"][:300]], axis=1)
    for exact in (False, True):
        offs, ids = f.neighbors_batch(locs, 10.0, exact)
        offs2, ag = f.neighbors_agents(locs, 10.0, exact)
        assert (offs == offs2).all() and (ids == ag["id"]).all() and len(ids) > 3000
        for k, src in (("x", "x"), ("y", "y"), ("ldx", "ldx"), ("ldy", "ldy")):
            assert (ag[k].view(np.uint32) == agents[src][ag["id"]].view(np.uint32)).all(), (exact, k)
    for q in range(50):     # single queries, growing and shrinking result sizes
        loc = (float(agents["x"][q]), float(agents["y"][q]))
        one = f.neighbors_agents([loc], 10.0 + q % 7, False)[1]["id"]
        assert (one == f.get_neighbors_within_relax_distance(loc, 10.0 + q % 7)).all()
    f.remove_object_location((3, 0.0, 0.0), (float(agents["x"][3]), float(agents["y"][3])))   # nothing in the log: no-op
    f.close()


def test_series_recorded_on_the_device_equals_reduce_after_every_step():
    """kg_field2d_run_boids_series (a plot! series without a host round trip per step): its rows are the
    reductions a step-by-step run reads back, bit for bit (deterministic f64 sums)"""
    n, w = 30000, 700.0
    agents = random_agents(n, w, w, seed=55)
    _, gp = both_params(exact=0, seed=8)
    fs = []
    for _ in range(2):
        f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=n)
        f.set_order(True)
        f.set_object_locations(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
        f.lazy_update()
        fs.append(f)
    gp.step = 0
    rows = fs[0].run_boids_series(gp, 24, every=3)
    assert rows.shape == (8, 8)
    keys = ("sum_x", "sum_y", "sum_ldx", "sum_ldy", "sum_speed", "sum_xx", "sum_yy")
    for r in range(8):
        gp.step = 3 * r
        fs[1].run_boids(gp, 3)
        red = fs[1].reduce()
        assert [red[k] for k in keys] == list(rows[r][:7]), r
    from krabmaga_b200 import plots
    plots.DATA.clear()
    kb.addplot("Flock", "step", "mean speed")
    gp.step = 24
    kb.plot_series("Flock", "speed", fs[0], gp, 10, 5, lambda red: red["sum_speed"] / red["n"], x0=24)
    pts = plots.DATA["Flock"].series["speed"]
    assert [x for x, _ in pts] == [29.0, 34.0] and all(abs(y - 0.7) < 1e-5 for _, y in pts)
    for f in fs:
        f.close()
