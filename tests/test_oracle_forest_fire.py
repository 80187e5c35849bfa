"""Row K (the Forest-Fire rule) pinned a second time: the oracle runs the rule through its restated
DenseNumberGrid2D API (get_value / set_value_location per live cell, lazy_update); here the same
rule is restated on whole numpy arrays, sharing no code with oracle/ — the reference ships no such
model, so two independent restatements agreeing is what pins the CUDA kernels' target.
The GPU suite compares K5 and the fused multi-step passes with the oracle on the same kind of
states (tests/test_gpu_grid.py)."""
import numpy as np
import pytest

import oracle_binding as ob

GREEN, BURNING, BURNED, NONE = 1, 2, 3, 0xFF


def numpy_step(v):
    """one synchronous step: a green cell with a burning Moore-8 neighbour ignites, a burning cell
    burns out, everything else (burned cells, empty cells) stays; no neighbours beyond the grid"""
    burning = v == BURNING
    p = np.pad(burning, 1)
    fire = np.zeros_like(burning)
    for dx in (0, 1, 2):
        for dy in (0, 1, 2):
            if (dx, dy) != (1, 1):
                fire |= p[dx:dx + v.shape[0], dy:dy + v.shape[1]]
    out = v.copy()
    out[(v == GREEN) & fire] = BURNING
    out[burning] = BURNED
    return out


@pytest.mark.parametrize("w,h", [(1, 1), (1, 16), (7, 7), (33, 50), (96, 64), (40, 300)])
def test_oracle_rule_equals_the_numpy_restatement_on_random_states(w, h):
    rng = np.random.default_rng(w * 1009 + h)
    cells = rng.choice(np.array([GREEN, BURNING, BURNED, NONE], np.uint8), size=(w, h), p=[0.7, 0.04, 0.06, 0.2])
    o = ob.ForestFire(w, h)
    o.load(cells)
    want = cells
    for step in range(1, 26):
        o.step(1)
        want = numpy_step(want)
        assert (o.dump() == want).all(), step


def test_oracle_init_and_front_speed():
    """the synthetic initial state of BASELINE config 4 (trees with probability 0.6, column x = 0
    burning) and what follows from the rule: None never changes, states only advance, the front
    moves at most one row per step"""
    w, h = 64, 128
    o = ob.ForestFire(w, h)
    o.init(0.6, 42)
    s0 = o.dump()
    assert set(np.unique(s0)) <= {GREEN, BURNING, NONE}
    assert ((s0[0] == BURNING) | (s0[0] == NONE)).all() and (s0[1:] != BURNING).all()
    assert 0.5 < (s0 != NONE).mean() < 0.7
    want = s0
    for step in range(1, 31):
        o.step(1)
        want = numpy_step(want)
    got = o.dump()
    assert (got == want).all()
    assert ((got == NONE) == (s0 == NONE)).all()
    live = s0 != NONE
    assert (got[live] >= s0[live]).all()
    assert (got[31:][live[31:]] == GREEN).all()
