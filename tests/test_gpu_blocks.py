"""2-D block decomposition (csrc/block.cu) against one GPU: with KG_ORDER_CANONICAL on both sides
every owned agent sees the same candidate sequence, so the blocks must reproduce the single field bit for bit —
across seams in both directions, block corners, the world's edges (clamped windows) and its wrap (migrants)."""
import numpy as np
import pytest

import krabmaga_b200 as kb
from krabmaga_b200 import _abi as abi
from krabmaga_b200 import blocks
from parity_util import NORTH_STAR_DISC, both_params, by_id, random_agents

pytestmark = pytest.mark.gpu


def devices_for(n):
    nd = abi.lib().kg_device_count()
    return [k % max(nd, 1) for k in range(n)]


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


def edge_cases(agents, w):
    agents["x"][:8] = [0.0, w - 1e-3, 1e-3, w / 2, w / 2 - 1e-3, w / 3, w, 2 * w / 3]
    agents["y"][:8] = [w - 1e-3, 0.0, w / 2, 1e-3, w / 2 + 1e-3, w, w / 3, 2 * w / 3 - 1e-3]
    return agents


@pytest.mark.parametrize("nbx,nby", [(1, 1), (2, 2), (3, 2), (1, 3), (4, 1), (2, 4)])
@pytest.mark.parametrize("exact", [0, 1])
def test_blocks_reproduce_single_gpu_bit_exact(nbx, nby, exact):
    n, w, nsteps = 20000, 600.0, 40
    agents = edge_cases(random_agents(n, w, w, seed=17 + nbx * 10 + nby), w)
    _, gp = both_params(exact=exact, seed=77, cohesion=1.2, avoidance=0.85, consistency=0.9, randomness=1.4,
                        momentum=0.95)
    want = single_gpu(agents, w, nsteps, gp)
    world = blocks.BlockWorld(w, w, NORTH_STAR_DISC, 10.0, nbx, nby, devices_for(nbx * nby), n, canonical_order=True,
                              slack=3.0)
    world.upload(agents)
    gp.step = 0
    world.run_boids(gp, nsteps)
    d = world.download()
    assert sorted(d["id"].tolist()) == list(range(n))          # every agent owned by exactly one block
    got = by_id(d)
    for k in want:
        bad = np.flatnonzero(got[k].view(np.uint32) != want[k].view(np.uint32))
        assert len(bad) == 0, f"{k}: {len(bad)} of {n} differ (ids {bad[:5]})"
    world.close()


def test_blocks_with_a_wide_window():
    """disc 3, radius 10: a 7 x 7 cell window, three halo cells around every block"""
    n, w, nsteps = 12000, 600.0, 12
    agents = edge_cases(random_agents(n, w, w, seed=5), w)
    for exact in (0, 1):
        _, gp = both_params(exact=exact, seed=3, cohesion=1.1, avoidance=0.9)
        want = single_gpu(agents, w, nsteps, gp, disc=3.0)
        world = blocks.BlockWorld(w, w, 3.0, 10.0, 3, 3, devices_for(9), n, canonical_order=True, slack=3.0)
        world.upload(agents)
        gp.step = 0
        world.run_boids(gp, nsteps)
        got = by_id(world.download())
        for k in want:
            assert (got[k].view(np.uint32) == want[k].view(np.uint32)).all(), (exact, k)
        world.close()


def test_block_geometry_and_ownership():
    w = 600.0
    world = blocks.BlockWorld(w, w, NORTH_STAR_DISC, 10.0, 3, 2, devices_for(6), 6000, slack=3.0)
    max_x = 90                                              # ceil(600 / 6.6666665)
    cols, rows = set(), set()
    for k in range(6):
        own, loc = world.cells(k)
        cols.add((own[0], own[1]))
        rows.add((own[2], own[3]))
        assert loc[0] == max(own[0] - 1, 0) and loc[1] == min(own[1] + 1, max_x + 1)
    assert sorted(cols) == [(0, 30), (30, 60), (60, 91)]    # the last block also owns the padding column
    assert sorted(rows) == [(0, 45), (45, 91)]
    agents = random_agents(6000, w, w, seed=2)
    world.upload(agents)
    held = sum(a for a, _ in world.counts())
    assert held > 6000                                      # ghosts are held on top of the owned agents
    per = world.download(per_block=True)
    assert sum(len(p["id"]) for p in per) == 6000
    for k, p in enumerate(per):
        own, _ = world.cells(k)
        cx = np.floor(p["x"] / np.float32(NORTH_STAR_DISC)).astype(int)
        cy = np.floor(p["y"] / np.float32(NORTH_STAR_DISC)).astype(int)
        assert ((cx >= own[0]) & (cx < own[1]) & (cy >= own[2]) & (cy < own[3])).all()
    world.close()


def test_blocks_refuse_what_they_cannot_do():
    with pytest.raises(kb.KgError):
        blocks.BlockWorld(60.0, 60.0, NORTH_STAR_DISC, 10.0, 8, 1, [0], 100)       # blocks thinner than the window
    world = blocks.BlockWorld(600.0, 600.0, NORTH_STAR_DISC, 10.0, 2, 2, devices_for(4), 100)
    a = random_agents(10, 600.0, 600.0, seed=1)
    a["id"][3] = a["id"][4]
    with pytest.raises(ValueError):
        world.upload(a)
    _, gp = both_params(radius=30.0)
    with pytest.raises(kb.KgError):
        world.step_boids(gp)                                                       # halo built for radius 10
    world.close()


@pytest.mark.parametrize("n", [1, 5, 60])
def test_blocks_with_empty_blocks_and_a_crowded_corner(n):
    """most blocks hold nobody; then everybody sits in one block's corner and walks out of it"""
    w, nsteps = 600.0, 25
    agents = random_agents(n, w, w, seed=n)
    for crowd in (False, True):
        if crowd:
            agents["x"] = (agents["x"] * np.float32(0.02) + np.float32(w / 3 - 3.0)).astype(np.float32)
            agents["y"] = (agents["y"] * np.float32(0.02) + np.float32(w / 2 - 3.0)).astype(np.float32)
        _, gp = both_params(exact=0, seed=5)
        want = single_gpu(agents, w, nsteps, gp)
        world = blocks.BlockWorld(w, w, NORTH_STAR_DISC, 10.0, 3, 2, devices_for(6), n, canonical_order=True, slack=8.0)
        world.upload(agents)
        gp.step = 0
        world.run_boids(gp, nsteps)
        got = by_id(world.download())
        for k in want:
            assert (got[k].view(np.uint32) == want[k].view(np.uint32)).all(), (crowd, k)
        world.close()
