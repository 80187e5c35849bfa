"""Run-time compiled closures (csrc/jit.cuh): `apply_to_all_values` with an arbitrary closure body
(dense_number_grid_2d.rs:155-195) and a grid rule given as an expression, checked against the oracle where
the oracle has the closure, against the shipped kernels, and against numpy restatements of the closure."""
import numpy as np
import pytest

import krabmaga_b200 as kb
import oracle_binding as ob
from krabmaga_b200 import GridOption
from krabmaga_b200 import _abi as abi

pytestmark = pytest.mark.gpu
NONE16 = 0xFFFF

FIRE = ("v == 1 ? (at(-1,-1)==2 || at(-1,0)==2 || at(-1,1)==2 || at(0,-1)==2 || at(0,1)==2 || at(1,-1)==2 || "
        "at(1,0)==2 || at(1,1)==2 ? 2 : 1) : (v == 2 ? 3 : v)")


def two_phase_grids(W, H, elem, seed):
    rng = np.random.default_rng(seed)
    o = ob.DenseNumberGrid2D(W, H)
    g = kb.DenseNumberGrid2D(W, H, elem_size=elem)
    for phase in range(2):      # phase 0 fills what becomes the read buffer, phase 1 the write buffer
        m = rng.random((W, H)) < 0.5
        xs, ys = np.nonzero(m)
        vals = rng.integers(0, 50, len(xs))
        for x, y, v in zip(xs, ys, vals):
            o.set_value_location(int(v), int(x), int(y))
        g.set_values(xs, ys, vals)
        if phase == 0:
            o.lazy_update()
            g.lazy_update()
    return o, g


@pytest.mark.parametrize("option", [GridOption.READ, GridOption.WRITE, GridOption.READWRITE])
@pytest.mark.parametrize("elem", [1, 2, 4])
def test_expression_closure_matches_the_oracle_closure(option, elem):
    """|v| v + 3 as the string "v + 3": all three GridOption arms over partially filled buffers"""
    o, g = two_phase_grids(13, 9, elem, int(option) * 10 + elem)
    o.apply_add(3, int(option))
    g.apply_to_all_values("v + 3", option)
    for unbuf in (False, True):
        want = o.dump(unbuffered=unbuf).astype(np.int64)
        got = g.download(unbuffered=unbuf).astype(np.int64)
        want[want == NONE16] = -1
        got[got == g.none] = -1
        assert (want == got).all()


def test_doc_example_and_a_closure_using_the_cell():
    """the reference's doc example `grid.apply_to_all_values(|x| x - 1, GridOption::READ)` (:152), then a
    closure no fixed family covers, restated in numpy"""
    W, H = 17, 11
    g = kb.DenseNumberGrid2D(W, H, elem_size=4)
    xs, ys = np.meshgrid(np.arange(W), np.arange(H), indexing="ij")
    g.set_values(xs.ravel(), ys.ravel(), (xs * H + ys + 5).ravel())
    g.lazy_update()
    g.apply_to_all_values("v - 1", GridOption.READ)
    assert (g.download() == xs * H + ys + 4).all()
    g.apply_to_all_values("v % 3 == 0 ? v * 7 + x : (v + 2 * y) % 1000", GridOption.READ)
    v = xs * H + ys + 4
    want = np.where(v % 3 == 0, v * 7 + xs, (v + 2 * ys) % 1000)
    assert (g.download() == want).all()
    g.close()


def test_a_closure_that_does_not_compile_and_one_that_returns_none():
    g = kb.DenseNumberGrid2D(4, 4)
    g.set_values([0, 1], [0, 1], [254, 7])
    g.lazy_update()
    with pytest.raises(kb.KgError) as e:
        g.apply_to_all_values("v +* nonsense", GridOption.READ)
    assert e.value.code == abi.KG_E_INVALID and "compile" in str(e.value)
    g.apply_to_all_values("v + 1", GridOption.READ)           # 254 + 1 is the reserved value
    with pytest.raises(kb.KgError) as e:
        g.sync()
    assert e.value.code == abi.KG_E_INVALID
    g.close()


@pytest.mark.parametrize("w,h", [(64, 64), (37, 48), (1, 16), (7, 7), (130, 200)])
def test_forest_fire_as_an_expression_equals_the_oracle_and_the_shipped_rule(w, h):
    o = ob.ForestFire(w, h)
    o.init(0.6, 42)
    g = kb.DenseNumberGrid2D(w, h)
    s = kb.DenseNumberGrid2D(w, h)
    g.init_forest_fire(0.6, 42)
    s.init_forest_fire(0.6, 42)
    for _ in range(25):
        o.step(1)
        g.step_stencil(FIRE)
        g.lazy_update()
        s.step_stencil()
        s.lazy_update()
        assert (g.download() == o.dump()).all() and (s.download() == o.dump()).all()


def test_expression_rule_writes_live_cells_only():
    """like the shipped rule: a value the model put into the write buffer at a None cell survives the step"""
    w, h = 32, 32
    cells = np.full((w, h), 0xFF, np.uint8)
    cells[3:20, 4:28] = 1
    cells[10, 10] = 2
    o = ob.ForestFire(w, h)
    o.load(cells)
    g = kb.DenseNumberGrid2D(w, h)
    g.upload(cells, unbuffered=True)
    g.lazy_update()
    g.set_value_location(3, (0, 0))
    g.step_stencil(FIRE)
    g.lazy_update()
    o.step(1)
    want = o.dump()
    want[0, 0] = 3
    assert (g.download() == want).all()


def test_game_of_life_rule_against_numpy():
    """a rule the library does not ship: Conway's Life on a bounded grid (0 dead, 1 alive), u16 cells"""
    w, h = 50, 70
    rng = np.random.default_rng(3)
    cells = (rng.random((w, h)) < 0.35).astype(np.uint16)
    g = kb.DenseNumberGrid2D(w, h, elem_size=2)
    g.upload(cells, unbuffered=True)
    g.lazy_update()
    nb = "+".join(f"(at({dx},{dy})==1)" for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0))
    rule = f"(({nb}) == 3 || (v == 1 && ({nb}) == 2)) ? 1 : 0"
    cur = cells.astype(np.int64)
    for _ in range(12):
        g.step_stencil(rule)
        g.lazy_update()
        p = np.pad(cur, 1)
        n = sum(p[1 + dx:1 + dx + w, 1 + dy:1 + dy + h] for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0))
        cur = ((n == 3) | ((cur == 1) & (n == 2))).astype(np.int64)
        assert (g.download() == cur).all()
    g.close()
