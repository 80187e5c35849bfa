"""Batched replicas (kg_batch_*, the device-side explore_parallel!) against single-field runs and
the oracle: every replica must be exactly the simulation it would have been on its own."""
import numpy as np
import pytest

import krabmaga_b200 as kb
import oracle_binding as ob
from krabmaga_b200 import _abi as abi
from parity_util import NORTH_STAR_DISC, by_id, random_agents

pytestmark = pytest.mark.gpu


def replica_by_id(d, r):
    n = d["id"].shape[1]
    out = {}
    for k in ("x", "y", "ldx", "ldy"):
        a = np.zeros(n, np.float32)
        a[d["id"][r]] = d[k][r]
        out[k] = a
    return out


def sweep_params(R, exact=0, radius=10.0):
    ps = []
    for r in range(R):
        ps.append(abi.boids_params(radius=radius, exact=exact, seed=42 + r, cohesion=1.0 + 0.25 * r,
                                   avoidance=1.0 - 0.1 * (r % 3), consistency=0.5 + 0.5 * (r % 2),
                                   momentum=1.0, randomness=1.0 + 0.125 * r))
    return ps


def oparams(p):
    return ob.boids_params(radius=p.radius, exact=p.exact_query, seed=p.seed, jump=p.jump,
                           cohesion=p.cohesion, avoidance=p.avoidance, randomness=p.randomness,
                           consistency=p.consistency, momentum=p.momentum)


@pytest.mark.parametrize("exact,radius,w", [(0, 10.0, 150.0), (1, 10.0, 150.0), (0, 14.0, 150.0),
                                            (0, 10.0, 24.0), (1, None, 150.0), (0, None, 150.0)])
def test_every_replica_matches_the_oracle_bit_for_bit(exact, radius, w):
    """canonical in-bag order on both sides; Philox init with each replica's own seed; the 24x24
    world is too small for the packed kernels and takes the generic path; radius None = a different
    radius per replica inside one window size (each replica has its own exact-query threshold)"""
    R, n, steps = 5, 1400 if w > 100 else 60, 12
    ps = sweep_params(R, exact, radius or 10.0)
    if radius is None:
        for r, rad in enumerate((7.0, 9.5, 10.0, 12.25, 13.3)):
            ps[r].radius = rad
    b = kb.FlockerBatch((w, w), n, R, NORTH_STAR_DISC, True, ps, canonical_order=True)
    b.init()
    b.run(steps)
    got = b.download(with_cells=True)
    b.close()
    for r in range(R):
        m = ob.Flockers(w, w, n, NORTH_STAR_DISC, True, oparams(ps[r]), canonical_order=True)
        m.init()
        m.step(steps)
        x, y, dx, dy = m.agents()
        g = replica_by_id(got, r)
        assert (np.sort(got["id"][r]) == np.arange(n)).all()
        for k, want in (("x", x), ("y", y), ("ldx", dx), ("ldy", dy)):
            assert (g[k].view(np.uint32) == want.view(np.uint32)).all(), (r, k)
        it = m.field1.iter_objects()
        assert (got["id"][r] == it["id"]).all()          # same iter_objects order
        assert (got["cell"][r] == it["cell"]).all()


def test_mixed_windows_fall_back_to_the_generic_kernel():
    """replicas whose radii give different window sizes cannot share the packed kernel"""
    R, n, w, steps = 4, 900, 120.0, 6
    ps = sweep_params(R)
    ps[1].radius = 14.0   # dd = 2
    ps[3].exact_query = 1
    b = kb.FlockerBatch((w, w), n, R, NORTH_STAR_DISC, True, ps, canonical_order=True)
    b.init()
    b.run(steps)
    got = b.download()
    b.close()
    for r in range(R):
        m = ob.Flockers(w, w, n, NORTH_STAR_DISC, True, oparams(ps[r]), canonical_order=True)
        m.init()
        m.step(steps)
        x, y, dx, dy = m.agents()
        g = replica_by_id(got, r)
        assert (g["x"].view(np.uint32) == x.view(np.uint32)).all(), r
        assert (g["ldy"].view(np.uint32) == dy.view(np.uint32)).all(), r


def test_uploaded_population_equals_single_field_runs():
    """host data in (ids need not be 0..n-1): each replica == the same agents on a kg_field2d"""
    R, n, w, steps = 3, 2000, 200.0, 8
    ps = sweep_params(R)
    pops = [random_agents(n, w, w, seed=900 + r) for r in range(R)]
    for r in range(R):
        pops[r]["id"] = (pops[r]["id"] * 3 + 7).astype(np.uint32)     # sparse, still unique
    cat = {k: np.concatenate([p[k] for p in pops]) for k in pops[0]}
    b = kb.FlockerBatch((w, w), n, R, NORTH_STAR_DISC, True, ps, canonical_order=True)
    b.upload(cat["id"], cat["x"], cat["y"], cat["ldx"], cat["ldy"])
    b.run(steps)
    got = b.download()
    b.close()
    for r in range(R):
        f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=n)
        f.set_order(True)
        f.set_object_locations(pops[r]["id"], pops[r]["x"], pops[r]["y"], pops[r]["ldx"], pops[r]["ldy"])
        f.lazy_update()
        p = ps[r]
        p.step = 0
        f.run_boids(p, steps)
        want = f.download()
        f.close()
        assert (got["id"][r] == want["id"]).all()
        for k in ("x", "y", "ldx", "ldy"):
            assert (got[k][r].view(np.uint32) == want[k].view(np.uint32)).all(), (r, k)


def test_out_of_world_upload_is_rejected():
    b = kb.FlockerBatch((50.0, 50.0), 4, 2, NORTH_STAR_DISC)
    ids = np.arange(8, dtype=np.uint32) % 4
    x = np.full(8, 10.0, np.float32)
    y = np.full(8, 10.0, np.float32)
    x[5] = 1e6
    with pytest.raises(kb.KgOutOfBounds):
        b.upload(ids, x, y, np.zeros(8, np.float32), np.zeros(8, np.float32))
    x[5] = 10.0
    b.upload(ids, x, y, np.zeros(8, np.float32), np.zeros(8, np.float32))
    b.run(2)
    assert b.download()["id"].shape == (2, 4)
    b.close()


def test_sweep_size_properties_256_replicas_of_16k():
    """config 5 shape (16 384 agents in 512x512 per replica), 256 replicas: conservation, replica
    isolation (identical seeds and weights give identical replicas, different seeds do not),
    |last_d| = JUMP"""
    R, n, w = 256, 16384, 512.0
    ps = [abi.boids_params(radius=10.0, exact=0, seed=42 + (r % 128)) for r in range(R)]
    b = kb.FlockerBatch((w, w), n, R, NORTH_STAR_DISC, True, ps, canonical_order=True)
    b.init()
    b.run(10)
    d = b.download()
    b.close()
    assert (np.sort(d["id"], axis=1) == np.arange(n)[None, :]).all()
    norm = np.sqrt(d["ldx"].astype(np.float64) ** 2 + d["ldy"].astype(np.float64) ** 2)
    assert np.abs(norm - 0.7).max() < 1e-5
    assert ((d["x"] >= 0) & (d["x"] <= w) & (d["y"] >= 0) & (d["y"] <= w)).all()
    for k in ("id", "x", "y", "ldx", "ldy"):
        assert (d[k][:128].view(np.uint32) == d[k][128:].view(np.uint32)).all(), k
    assert (d["x"][0] != d["x"][1]).any()


def test_explore_parallel_rows_and_values():
    """explore_parallel!: n_conf * rep_conf rows in run order, conf_num = run / rep_conf,
    conf_rep = run % rep_conf (model_exploration.rs:387-420); each row's output equals the same
    simulation run alone"""
    inputs = {"cohesion": [0.5, 1.0, 2.0], "avoidance": [1.0, 3.0]}
    rows = kb.explore_parallel(15, 2, (120.0, 120.0), 900, NORTH_STAR_DISC, inputs,
                               mode=kb.ExploreMode.Exaustive, max_replicas_per_batch=5, canonical_order=True)
    assert len(rows) == 12
    assert [(r["conf_num"], r["conf_rep"]) for r in rows] == [(i, k) for i in range(6) for k in range(2)]
    assert [(r["cohesion"], r["avoidance"]) for r in rows[::2]] == \
        [(0.5, 1.0), (0.5, 3.0), (1.0, 1.0), (1.0, 3.0), (2.0, 1.0), (2.0, 3.0)]
    assert all(r["run_duration"] > 0 and r["step_per_sec"] > 0 for r in rows)
    seq = kb.explore_sequential(15, 2, (120.0, 120.0), 900, NORTH_STAR_DISC, inputs,
                                mode=kb.ExploreMode.Exaustive, canonical_order=True)
    # canonical in-bag order: a run's result does not depend on which batch it ran in
    assert [r["polarisation"] for r in rows] == [r["polarisation"] for r in seq]
    assert len({r["polarisation"] for r in rows}) == 12     # and the runs really differ
    matched = kb.explore_parallel(5, 1, (120.0, 120.0), 300, NORTH_STAR_DISC,
                                  {"cohesion": [1.0, 2.0], "seed": [7, 9]}, mode=kb.ExploreMode.Matched)
    assert [(r["cohesion"], r["seed"]) for r in matched] == [(1.0, 7), (2.0, 9)]


def test_device_reductions_match_the_downloaded_population():
    """kg_batch_reduce / kg_field2d_reduce: output columns from device-side sums"""
    R, n, w = 24, 3000, 220.0
    ps = [abi.boids_params(radius=10.0, exact=0, seed=100 + r, cohesion=1.0 + 0.1 * r) for r in range(R)]
    b = kb.FlockerBatch((w, w), n, R, NORTH_STAR_DISC, True, ps)
    b.init()
    b.run(12)
    red, d = b.reduce(), b.download()
    for key, col in (("sum_x", "x"), ("sum_y", "y"), ("sum_ldx", "ldx"), ("sum_ldy", "ldy")):
        want = d[col].astype(np.float64).sum(axis=1)
        assert np.allclose(red[key], want, rtol=1e-12, atol=1e-9), key
    speed = np.sqrt(d["ldx"].astype(np.float64) ** 2 + d["ldy"].astype(np.float64) ** 2).sum(axis=1)
    assert np.allclose(red["sum_speed"], speed, rtol=1e-12)
    assert np.allclose(red["sum_xx"], (d["x"].astype(np.float64) ** 2).sum(axis=1), rtol=1e-12)
    again = b.reduce()
    assert all((again[k] == red[k]).all() for k in red)          # deterministic
    pol = kb.explore.default_outputs(red)["polarisation"]
    vx, vy = d["ldx"].mean(axis=1, dtype=np.float64), d["ldy"].mean(axis=1, dtype=np.float64)
    assert np.allclose(pol, np.sqrt(vx * vx + vy * vy) / 0.7, rtol=1e-9)
    b.close()
    f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=50000)
    f.init_flockers(50000, 3)
    f.lazy_update()
    f.run_boids(abi.boids_params(radius=10.0, exact=0, seed=3), 5)
    r1, dd = f.reduce(), f.download(with_cells=False)
    assert r1["n"] == 50000 and np.isclose(r1["sum_x"], dd["x"].astype(np.float64).sum(), rtol=1e-12)
    assert np.isclose(r1["sum_ldy"], dd["ldy"].astype(np.float64).sum(), rtol=1e-10, atol=1e-9)
    f.close()


def test_sweep_over_world_and_population_equals_separate_runs():
    """`dim` / `initial_flockers` as swept inputs (explore_parallel! builds an arbitrary state per
    configuration, model_exploration.rs:387-410): grouped into per-world batches, same rows as one
    run at a time"""
    inputs = {"cohesion": [0.8, 1.2, 1.0], "initial_flockers": [700, 1100, 700],
              "dim": [(110.0, 110.0), (140.0, 140.0), (110.0, 110.0)]}
    rows = kb.explore_parallel(10, 2, (1.0, 1.0), 1, NORTH_STAR_DISC, inputs, mode=kb.ExploreMode.Matched,
                               canonical_order=True)
    seq = kb.explore_sequential(10, 2, (1.0, 1.0), 1, NORTH_STAR_DISC, inputs, mode=kb.ExploreMode.Matched,
                                canonical_order=True)
    assert [(r["conf_num"], r["conf_rep"], r["initial_flockers"]) for r in rows] == \
        [(i, k, n) for i, n in enumerate([700, 1100, 700]) for k in range(2)]
    assert [r["polarisation"] for r in rows] == [r["polarisation"] for r in seq]
    assert len({r["polarisation"] for r in rows}) == 6
