"""GPU tests of the strip decomposition (halo exchange + migration over peer memory).

The union of the strips must reproduce the single-GPU result: cell assignments / neighbour sets
exact, and — with KG_ORDER_CANONICAL on both sides — positions bit for bit, because every agent
sees exactly the same candidate sequence.  Strips are placed round-robin on the visible devices;
on a one-GPU box they share the device, which exercises the same kernels and inbox protocol."""
import numpy as np
import pytest

import krabmaga_b200 as kb
from krabmaga_b200 import strips
from parity_util import NORTH_STAR_DISC, both_params, by_id, random_agents

pytestmark = pytest.mark.gpu


def devices_for(n):
    nd = kb._abi.lib().kg_device_count()
    return [r % nd for r in range(n)]


def single_gpu(agents, w, nsteps, gp, disc=NORTH_STAR_DISC):
    f = kb.Field2D(w, w, disc, True, capacity=len(agents["id"]))
    f.set_order(True)
    f.set_object_locations(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    f.lazy_update()
    gp.step = 0
    f.run_boids(gp, nsteps)
    out = by_id(f.download())
    f.close()
    return out


@pytest.mark.parametrize("nranks", [1, 2, 3, 4])
def test_strips_reproduce_single_gpu_bit_exact(nranks):
    n, w, nsteps = 20000, 600.0, 60
    agents = random_agents(n, w, w, seed=5)
    # seed agents right at strip boundaries and at the wrap-around seam
    agents["x"][:6] = [0.0, w - 1e-3, 1e-3, w / 2, w / 2 - 1e-3, w / 3]
    _, gp = both_params(exact=0, seed=77, cohesion=1.2, avoidance=0.85, consistency=0.9, randomness=1.4,
                        momentum=0.95)   # non-unit weights: see test_every_k4_variant_...
    want = single_gpu(agents, w, nsteps, gp)
    world = strips.StripWorld(w, w, NORTH_STAR_DISC, 10.0, devices_for(nranks), n,
                              canonical_order=True, slack=3.0)
    world.upload(agents)
    gp.step = 0
    world.run_boids(gp, nsteps)
    d = world.download()
    assert len(d["id"]) == n and (np.sort(d["id"]) == np.arange(n)).all()
    got = by_id(d)
    for k in want:
        bad = np.flatnonzero(got[k].view(np.uint32) != want[k].view(np.uint32))
        assert len(bad) == 0, f"{k}: {len(bad)} of {n} differ (ids {bad[:5]})"
    st = world.stats()
    assert sum(s["n_owned"] for s in st) == n
    if nranks > 1:
        assert sum(s["migrants_in"] for s in st) == sum(s["migrants_out"] for s in st) > 0
        assert all(s["halo_left"] > 0 for s in st[1:]) and all(s["halo_right"] > 0 for s in st[:-1])
        assert st[0]["halo_left"] == 0 and st[-1]["halo_right"] == 0    # clamped window: no wrap halo
    world.close()


@pytest.mark.parametrize("nranks", [2, 3])
def test_strips_exact_distance_query_bit_exact(nranks):
    """get_neighbors_within_distance — the query the reference's own fixture calls (bird.rs:41) —
    across strip seams: check_circle's cell classification and the per-element distance test must
    see ghosts exactly as one GPU sees its own agents."""
    n, w, nsteps = 20000, 600.0, 40
    agents = random_agents(n, w, w, seed=9)
    agents["x"][:6] = [0.0, w - 1e-3, 1e-3, w / 2, w / 2 - 1e-3, w / 3]
    _, gp = both_params(exact=1, seed=31, cohesion=1.1, avoidance=0.9, consistency=0.8, randomness=1.3,
                        momentum=0.97)
    want = single_gpu(agents, w, nsteps, gp)
    world = strips.StripWorld(w, w, NORTH_STAR_DISC, 10.0, devices_for(nranks), n,
                              canonical_order=True, slack=3.0)
    world.upload(agents)
    gp.step = 0
    world.run_boids(gp, nsteps)
    got = by_id(world.download())
    for k in want:
        bad = np.flatnonzero(got[k].view(np.uint32) != want[k].view(np.uint32))
        assert len(bad) == 0, f"{k}: {len(bad)} of {n} differ (ids {bad[:5]})"
    world.close()


def test_strips_exact_query_on_wide_windows_bit_exact():
    """a 7 x 7 cell window (disc 3, radius 10: three halo columns per side), exact and relaxed query"""
    n, w, nsteps = 12000, 600.0, 12
    agents = random_agents(n, w, w, seed=13)
    for exact in (1, 0):
        _, gp = both_params(exact=exact, seed=3, cohesion=1.1, avoidance=0.9)
        f = kb.Field2D(w, w, 3.0, True, capacity=n)
        f.set_order(True)
        f.set_object_locations(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
        f.lazy_update()
        gp.step = 0
        f.run_boids(gp, nsteps)
        want = by_id(f.download())
        f.close()
        world = strips.StripWorld(w, w, 3.0, 10.0, devices_for(3), n, canonical_order=True, slack=3.0)
        world.upload(agents)
        gp.step = 0
        world.run_boids(gp, nsteps)
        got = by_id(world.download())
        for k in want:
            bad = np.flatnonzero(got[k].view(np.uint32) != want[k].view(np.uint32))
            assert len(bad) == 0, f"exact={exact} {k}: {len(bad)} of {n} differ (ids {bad[:5]})"
        world.close()


def test_strips_philox_init_equals_single_gpu():
    n, w = 30000, 700.0
    _, gp = both_params(exact=0, seed=42)
    f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=n)
    f.set_order(True)
    f.init_flockers(n, 42)
    f.lazy_update()
    f.run_boids(gp, 25)
    want = by_id(f.download())
    world = strips.StripWorld(w, w, NORTH_STAR_DISC, 10.0, devices_for(2), n, canonical_order=True,
                              slack=3.0)
    world.init_flockers(n, 42)
    world.run_boids(gp, 25)
    got = by_id(world.download())
    for k in want:
        assert (got[k].view(np.uint32) == want[k].view(np.uint32)).all()
    world.close()


def test_strip_owned_agents_stay_inside_their_columns():
    n, w = 20000, 600.0
    world = strips.StripWorld(w, w, NORTH_STAR_DISC, 10.0, devices_for(3), n, slack=3.0)
    world.init_flockers(n, 1)
    _, gp = both_params(exact=0, seed=1)
    world.run_boids(gp, 40)
    parts = strips.partition(w, w, NORTH_STAR_DISC, 3)
    for s, (x0, x1) in zip(world.strips, parts):
        d = s.download()
        cols = np.floor(d["x"] / np.float32(NORTH_STAR_DISC)).astype(int)
        assert ((cols >= x0) & (cols < x1)).all()
    world.close()


def test_strip_overflow_is_reported():
    n, w = 20000, 600.0
    s = strips.StripField2D(w, w, NORTH_STAR_DISC, 10.0, 0, 1, capacity=1000, halo_capacity=64,
                            migrate_capacity=64)
    with pytest.raises(kb.KgError) as e:
        s.init_flockers(n, 1)
    assert e.value.code == kb._abi.KG_E_CAPACITY
    s.close()


def test_strip_rejects_unsupported_geometry():
    with pytest.raises(kb.KgError):
        strips.StripField2D(10.0, 10.0, 0.5, 10.0, 0, 2, 100, 100, 100)     # window >= half the world
    with pytest.raises(kb.KgError):
        strips.StripField2D(400.0, 400.0, NORTH_STAR_DISC, 10.0, 0, 2, 100, 100, 100, toroidal=False)
