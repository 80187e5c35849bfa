"""GPU parity tests for DenseNumberGrid2D and the Forest-Fire stencil (integer states: bit-exact)."""
import numpy as np
import pytest

import krabmaga_b200 as kb
import oracle_binding as ob
from krabmaga_b200 import GridOption
from krabmaga_b200 import _abi as abi

pytestmark = pytest.mark.gpu
NONE16 = 0xFFFF


def test_dense_number_grid_2d_apply_kat():
    """tests/engine/dense_number_grid_2d.rs:31-89 (T = u16)"""
    W = H = 10
    g = kb.DenseNumberGrid2D(W, H, elem_size=2)
    xs, ys = np.meshgrid(np.arange(W), np.arange(H), indexing="ij")
    g.set_values(xs.ravel(), ys.ravel(), np.zeros(W * H))
    g.lazy_update()
    g.apply_to_all_values(("const", 1), GridOption.WRITE)
    g.lazy_update()
    assert (g.download() == 1).all()
    g.apply_to_all_values(("add", 1), GridOption.READWRITE)
    g.lazy_update()
    g.apply_to_all_values(("add", 1), GridOption.READ)
    for i in range(W):
        for j in range(H):
            assert g.get_value((i, j)) == 3
    g.set_values(xs.ravel(), ys.ravel(), (xs * ys).ravel())
    seen = []
    g.iter_values_unbuffered(lambda loc, v: seen.append(v == loc.x * loc.y == g.get_value_unbuffered(loc)))
    assert len(seen) == W * H and all(seen)
    g.lazy_update()
    seen = []
    g.iter_values(lambda loc, v: seen.append(v == loc.x * loc.y == g.get_value(loc)))
    assert len(seen) == W * H and all(seen)


@pytest.mark.parametrize("i,j", [(1, 1), (3, 7), (9, 9), (9, 1)])
def test_dense_number_grid_2d_bags_kat(i, j):
    """tests/engine/dense_number_grid_2d.rs:97-165"""
    W = H = 10
    g = kb.DenseNumberGrid2D(W, H, elem_size=2)
    assert len(g.get_empty_bags()) == W * H == g.num_empty_bags()
    assert g.get_random_empty_bag() is not None
    loc = (4, 2)
    g.set_value_location(10, loc)
    assert g.get_value_unbuffered(loc) == 10
    g.remove_value_location(loc)
    assert g.get_value_unbuffered(loc) is None
    g.set_value_location(10, loc)
    g.update()
    assert g.num_empty_bags() == W * H - 1
    xs, ys = np.meshgrid(np.arange(W), np.arange(H), indexing="ij")
    g.set_values(xs.ravel(), ys.ravel(), np.zeros(W * H))
    assert g.get_location_unbuffered(0) == (0, 0)
    g.set_value_location(5, (i, j))
    assert g.get_location_unbuffered(5) == (i, j)
    assert g.get_location_unbuffered(6) is None
    g.lazy_update()
    assert g.get_location(0) == (0, 0)
    assert g.get_location(5) == (i, j)
    assert g.get_location(6) is None
    assert g.num_empty_bags() == 0


def test_lazy_update_clears_unwritten_cells_and_unbuffered_view():
    """dense_number_grid_2d.rs:537-545"""
    g = kb.DenseNumberGrid2D(4, 3, elem_size=2)
    g.set_value_location(7, (1, 2))
    g.lazy_update()
    assert g.get_value((1, 2)) == 7
    assert g.get_value_unbuffered((1, 2)) is None     # the new write buffer reads as all-None
    assert (g.download(unbuffered=True) == NONE16).all()
    g.lazy_update()
    assert g.get_value((1, 2)) is None and g.num_empty_bags() == 12


def test_two_swaps_in_a_row_leave_an_all_none_read_buffer():
    """tests/engine/dense_number_grid_2d.rs:160-165 — nothing reads or writes the write buffer
    between the swaps, so the deferred clear must still happen."""
    g = kb.DenseNumberGrid2D(10, 10)
    g.set_values(np.repeat(np.arange(10), 10), np.tile(np.arange(10), 10), np.zeros(100))
    g.lazy_update()
    assert g.num_empty_bags() == 0
    g.lazy_update()
    assert g.num_empty_bags() == 100
    assert (g.download() == 0xFF).all()


def test_grid_out_of_bounds():
    g = kb.DenseNumberGrid2D(4, 3)
    with pytest.raises(kb.KgOutOfBounds):
        g.set_value_location(1, (4, 3))
    with pytest.raises(kb.KgOutOfBounds):
        g.get_value((-1, 0))
    with pytest.raises(kb.KgError):
        kb.DenseNumberGrid2D(1 << 16, 1 << 16)  # i32 overflow of width*height


def test_option_none_sentinel_is_guarded():
    """Option<T> is one reserved value of T: storing Some(none) or computing it is an error, and
    get_location never matches empty cells (dense_number_grid_2d.rs:222 `elem.is_some() && ...`)."""
    g = kb.DenseNumberGrid2D(4, 4)                     # u8, none = 0xFF
    with pytest.raises(kb.KgError) as e:
        g.set_values([1], [1], [0xFF])
    assert e.value.code == abi.KG_E_INVALID
    with pytest.raises(kb.KgError):
        g.apply_to_all_values(("const", 0xFF), kb.GridOption.READ)
    g.set_values([0, 1], [0, 1], [254, 7])
    g.lazy_update()
    assert g.get_location(0xFF) is None and g.get_location(254) == (0, 0)
    g.apply_to_all_values(("add", 1), kb.GridOption.READ)   # 254 + 1 would become "None"
    with pytest.raises(kb.KgError) as e:
        g.sync()
    assert e.value.code == abi.KG_E_INVALID
    g.sync()                                                # the flag is consumed
    g.close()


@pytest.mark.parametrize("option", [GridOption.READ, GridOption.WRITE, GridOption.READWRITE])
@pytest.mark.parametrize("elem", [1, 2, 4])
def test_apply_matches_oracle_with_partial_buffers(option, elem):
    """apply_to_all_values over grids where read and write hold different sparse cells"""
    W, H = 13, 9
    rng = np.random.default_rng(int(option) * 10 + elem)
    o = ob.DenseNumberGrid2D(W, H)
    g = kb.DenseNumberGrid2D(W, H, elem_size=elem)
    none = g.none
    for phase in range(2):  # phase 0 fills what becomes the read buffer, phase 1 the write buffer
        m = rng.random((W, H)) < 0.5
        xs, ys = np.nonzero(m)
        vals = rng.integers(0, 50, len(xs))
        for x, y, v in zip(xs, ys, vals):
            o.set_value_location(int(v), int(x), int(y))
        g.set_values(xs, ys, vals)
        if phase == 0:
            o.lazy_update(); g.lazy_update()
    o.apply_add(3, int(option)); g.apply_to_all_values(("add", 3), option)
    for unbuf in (False, True):
        want = o.dump(unbuffered=unbuf).astype(np.int64)
        got = g.download(unbuffered=unbuf).astype(np.int64)
        want[want == NONE16] = -1
        got[got == none] = -1
        assert (want == got).all()


FF_SHAPES = [(64, 64), (96, 2048), (37, 48), (130, 4096 + 16), (33, 50), (1, 16), (16, 1), (7, 7),
             (200, 512)]


@pytest.mark.parametrize("w,h", FF_SHAPES)
def test_forest_fire_bit_exact(w, h):
    """K5 against the oracle's rule written through the DenseNumberGrid2D API, incl. shapes that
    take the 16-byte fast path (h % 16 == 0) and the generic path"""
    o = ob.ForestFire(w, h)
    o.init(0.6, 42)
    g = kb.DenseNumberGrid2D(w, h, elem_size=1)
    g.init_forest_fire(0.6, 42)
    assert (g.download() == o.dump()).all()       # Philox init identical
    for steps in (1, 1, 3, 10, 50):
        o.step(steps)
        g.run_stencil(steps)
        assert (g.download() == o.dump()).all()
    assert (g.download(unbuffered=True) == 0xFF).all()


def test_forest_fire_step_then_lazy_update_equals_run():
    w, h = 48, 64
    a = kb.DenseNumberGrid2D(w, h)
    b = kb.DenseNumberGrid2D(w, h)
    a.init_forest_fire(0.7, 3)
    b.init_forest_fire(0.7, 3)
    for _ in range(9):
        a.step_stencil()
        a.lazy_update()
    b.run_stencil(9)
    assert (a.download() == b.download()).all()


def test_forest_fire_respects_values_already_in_write_buffer():
    """step_stencil writes only live cells (set_value_location per live cell); a value the model put
    into the write buffer at a None cell survives until the swap"""
    w, h = 32, 32
    cells = np.full((w, h), 0xFF, np.uint8)
    cells[3:20, 4:28] = 1
    cells[10, 10] = 2
    o = ob.ForestFire(w, h)
    o.load(cells)
    g = kb.DenseNumberGrid2D(w, h)
    g.upload(cells, unbuffered=True)
    g.lazy_update()
    g.set_value_location(3, (0, 0))               # None in the read buffer
    g.step_stencil()
    g.lazy_update()
    o.step(1)
    want = o.dump()
    want[0, 0] = 3
    assert (g.download() == want).all()


@pytest.mark.parametrize("w,h", [(300, 4096), (65, 960), (130, 496), (2, 16), (64, 480), (67, 1936),
                                 (129, 32), (193, 8192 + 480)])
def test_forest_fire_two_steps_per_pass_equals_single_steps(w, h):
    """run_stencil advances runs of 8 / 4 / 2 steps with the fused bit-plane kernel (read step t, write
    step t+T, the steps between in registers); step_stencil + lazy_update never fuses.  Random states with fire everywhere,
    so every warp seam (960 cells of y), block seam and row-tile seam carries activity."""
    rng = np.random.default_rng(w * 100003 + h)
    cells = rng.choice(np.array([1, 2, 3, 0xFF], np.uint8), size=(w, h), p=[0.72, 0.03, 0.05, 0.2])
    o = ob.ForestFire(w, h)
    o.load(cells)
    a = kb.DenseNumberGrid2D(w, h)
    b = kb.DenseNumberGrid2D(w, h)
    for g in (a, b):
        g.upload(cells, unbuffered=True)
        g.lazy_update()
    done = 0
    for steps in (2, 3, 8, 1, 13):
        a.run_stencil(steps)
        for _ in range(steps):
            b.step_stencil()
            b.lazy_update()
        done += steps
        ga = a.download()
        assert (ga == b.download()).all(), f"after {done} steps"
        o.step(steps)
        assert (ga == o.dump()).all(), f"oracle, after {done} steps"
        assert (a.download(unbuffered=True) == 0xFF).all()


def test_forest_fire_fused_passes_on_arbitrary_bytes():
    """cells outside the model's alphabet (1, 2, 3, 0xFF): the fused kernel steps bit planes, K5 adds
    inside bytes — both only ever look at and move bits 0 and 1 and must leave the upper six alone"""
    w, h = 150, 1984
    rng = np.random.default_rng(11)
    cells = rng.integers(0, 256, size=(w, h), dtype=np.uint8)
    a = kb.DenseNumberGrid2D(w, h)
    b = kb.DenseNumberGrid2D(w, h)
    for g in (a, b):
        g.upload(cells, unbuffered=True)
        g.lazy_update()
    a.run_stencil(14)                             # 8 + 4 + 2
    for _ in range(14):
        b.step_stencil()
        b.lazy_update()
    assert (a.download() == b.download()).all()


@pytest.mark.parametrize("elem", [2, 4])
def test_forest_fire_wider_elements(elem):
    w, h = 40, 24
    o = ob.ForestFire(w, h)
    o.init(0.6, 9)
    g = kb.DenseNumberGrid2D(w, h, elem_size=elem)
    g.init_forest_fire(0.6, 9)
    o.step(12)
    g.run_stencil(12)
    got = g.download().astype(np.int64)
    got[got == g.none] = 0xFF
    assert (got == o.dump()).all()


def test_forest_fire_full_size_properties():
    """BASELINE config 4 geometry at reduced step count: monotone state, fire front speed 1"""
    w = h = 8192
    g = kb.DenseNumberGrid2D(w, h)
    g.init_forest_fire(0.6, 42)
    a = g.download().copy()
    steps = 25
    g.run_stencil(steps)
    b = g.download()
    assert ((a == 0xFF) == (b == 0xFF)).all()                     # None never changes
    live = a != 0xFF
    assert (b[live] >= a[live]).all()                             # GREEN < BURNING < BURNED
    assert (b[steps + 1:, :][live[steps + 1:, :]] == 1).all()     # front moves <= 1 row per step
    assert (b[0, :][live[0, :]] == 3).all()
    # one more step on a 512-row slab equals the oracle on that slab (top rows only: fire is there)
    slab = b[:64, :512].copy()
    o = ob.ForestFire(64, 512)
    o.load(slab)
    o.step(1)
    g2 = kb.DenseNumberGrid2D(64, 512)
    g2.upload(slab, unbuffered=True)
    g2.lazy_update()
    g2.run_stencil(1)
    assert (g2.download() == o.dump()).all()
