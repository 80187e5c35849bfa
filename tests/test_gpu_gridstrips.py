"""Row-strip decomposition of the Forest-Fire grid: G strips must reproduce the single-GPU grid
and the oracle bit for bit (integer states, BASELINE config 4's criterion)."""
import numpy as np
import pytest

import krabmaga_b200 as kb
import oracle_binding as ob
from krabmaga_b200 import gridstrips

pytestmark = pytest.mark.gpu


def devices_for(n):
    import torch
    have = max(1, torch.cuda.device_count())
    return [r % have for r in range(n)]


@pytest.mark.parametrize("w,h,G", [(96, 64, 2), (130, 48, 3), (64, 2064, 4), (9, 16, 8), (300, 4112, 5)])
def test_strips_equal_oracle(w, h, G):
    """ragged row counts, rows narrower than a 2048-cell span and wider, more strips than 64-row
    tiles, one-row strips"""
    steps = 37
    o = ob.ForestFire(w, h)
    o.init(0.6, 42)
    world = gridstrips.GridStripWorld(w, h, devices_for(G))
    world.init_forest_fire(0.6, 42)
    assert (world.download() == o.dump()).all()
    done = 0
    for upto in (1, 2, 11, steps):
        world.run_stencil(upto - done)
        o.step(upto - done)
        done = upto
        assert (world.download() == o.dump()).all(), upto
    world.close()


def test_uploaded_grid_and_reupload():
    """host data in; a second upload mid-run must not see stale halo rows"""
    w, h, G = 80, 96, 4
    rng = np.random.default_rng(3)
    cells = rng.choice(np.array([1, 1, 1, 2, 3, 0xFF], np.uint8), size=(w, h))
    world = gridstrips.GridStripWorld(w, h, devices_for(G))
    single = kb.DenseNumberGrid2D(w, h)
    for trial in range(2):
        world.upload(cells)
        single.upload(cells, unbuffered=True)
        single.lazy_update()
        for _ in range(3):
            world.run_stencil(5)
            single.run_stencil(5)
            assert (world.download() == single.download()).all()
        cells = np.roll(cells, 7, axis=0)   # a different state for the second round
    world.close()


def test_fire_front_crosses_every_strip_boundary():
    """full forest, left column burning: the front advances one row per step through all strips"""
    w, h, G = 64, 32, 4
    cells = np.full((w, h), 1, np.uint8)
    cells[0, :] = 2
    world = gridstrips.GridStripWorld(w, h, devices_for(G))
    world.upload(cells)
    world.run_stencil(40)
    got = world.download()
    assert (got[:40] == 3).all() and (got[40] == 2).all() and (got[41:] == 1).all()
    world.close()


def test_full_size_strip_properties():
    """8192 x 8192 over 4 strips == single grid (checksum of the whole state after 30 steps)"""
    w = h = 8192
    world = gridstrips.GridStripWorld(w, h, devices_for(4))
    world.init_forest_fire(0.6, 42)
    single = kb.DenseNumberGrid2D(w, h)
    single.init_forest_fire(0.6, 42)
    world.run_stencil(30)
    single.run_stencil(30)
    a, b = world.download(), single.download()
    assert (a == b).all()
    assert (a == 3).sum() > 0 and (a == 2).sum() > 0
    world.close()


@pytest.mark.parametrize("w,h,G", [(130, 960, 2), (131, 496, 2), (64, 2064, 4), (12, 32, 6), (9, 16, 8),
                                   (260, 1936, 3), (129, 480, 1)])
def test_two_steps_per_pass_over_strips_equal_single_steps(w, h, G):
    """run_stencil advances runs of 8 / 4 / 2 steps with the fused kernel (T halo rows per side, eight rows
    pushed per side by every pass); the single grid is stepped one step at a time, which never fuses.  Random
    states with fire everywhere; strips of 65 rows (last row tile would be one row), two-row and
    one-row strips (no fusing), odd step counts (fused passes followed by a single step)."""
    rng = np.random.default_rng(w * 7919 + h + G)
    cells = rng.choice(np.array([1, 2, 3, 0xFF], np.uint8), size=(w, h), p=[0.72, 0.03, 0.05, 0.2])
    world = gridstrips.GridStripWorld(w, h, devices_for(G))
    world.upload(cells)
    single = kb.DenseNumberGrid2D(w, h)
    single.upload(cells, unbuffered=True)
    single.lazy_update()
    done = 0
    for steps in (2, 3, 8, 1, 13):
        world.run_stencil(steps)
        for _ in range(steps):
            single.step_stencil()
            single.lazy_update()
        done += steps
        assert (world.download() == single.download()).all(), f"after {done} steps"
    world.close()
