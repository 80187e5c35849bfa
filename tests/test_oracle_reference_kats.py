"""Pins the oracle against the reference's own known-answer tests.

Each test is the reference test of the same name restated over the oracle binding; the
reference file:line it follows is in the docstring.  RNG draws in the reference tests
(`rand::rng()`) are replaced by sweeping every value the draw can take.
"""
import numpy as np
import pytest

import oracle_binding as ob

WIDTH, HEIGHT, DISC, TOROIDAL = 10.0, 10.0, 0.5, True  # tests/model/flockers/state.rs:11-14
NUM_AGENT = 10


def fixture_model(n, **kw):
    # the fixture's Bird::step uses the exact query with radius 10 (bird.rs:41)
    return ob.Flockers(WIDTH, HEIGHT, n, DISC, TOROIDAL, ob.boids_params(radius=10.0, exact=1), **kw)


def test_field_2d_single_step():
    """tests/engine/field_2d.rs:31-50"""
    m = fixture_model(NUM_AGENT)
    m.init()
    m.step(1)
    assert m.field1.nagents == NUM_AGENT
    assert len(m.field1.get_neighbors_within_distance(5.0, 5.0, 10.0)) == NUM_AGENT
    assert len(m.field1.get_neighbors_within_relax_distance(5.0, 5.0, 10.0)) == NUM_AGENT


@pytest.mark.parametrize("fly", [5.0, 6.0, 7.0, 8.0, 9.0])
def test_field_2d_neighbors(fly):
    """tests/engine/field_2d.rs:58-117 (fly = rng.random_range(5..10))"""
    f = ob.Field2D(WIDTH, HEIGHT, DISC, TOROIDAL)
    f.set_object_location(1, 0.0, 0.0)
    f.set_object_location(2, 0.0, 0.0)
    f.lazy_update()
    assert f.nagents == 2
    assert len(f.get_neighbors_within_distance(5.0, 5.0, 1.0)) == 0
    assert len(f.get_neighbors_within_relax_distance(5.0, 5.0, 1.0)) == 0
    f.set_object_location(1, fly, fly)
    f.set_object_location(2, 0.0, 0.0)
    f.lazy_update()
    assert list(f.get_neighbors_within_distance(fly, fly, 1.0)) == [1]
    assert list(f.get_neighbors_within_distance(0.0, 0.0, 1.0)) == [2]
    assert sorted(f.get_neighbors_within_distance(5.0, 5.0, 10.0)) == [1, 2]
    assert sorted(f.get_neighbors_within_relax_distance(5.0, 5.0, 10.0)) == [1, 2]


def test_field_2d_gets():
    """tests/engine/field_2d.rs:125-178"""
    f = ob.Field2D(WIDTH, HEIGHT, DISC, TOROIDAL)
    f.set_object_location(1, 0.0, 0.0)
    f.set_object_location(2, 5.0, 5.0)
    f.set_object_location(3, 5.0, 5.0)
    f.lazy_update()
    assert f.nagents == 3
    assert sorted(f.get_objects(5.0, 5.0)) == [2, 3]
    assert len(f.get_objects(10.0, 0.0)) == 0  # the +1 padding column exists (:150-151)
    assert f.num_objects_at_location(5.0, 5.0) == 2
    assert f.num_objects_at_location(0.0, 0.0) == 1
    f.set_object_location(4, 0.0, 0.0)
    assert len(f.get_objects_unbuffered(0.0, 0.0)) == 1
    assert len(f.get_objects(0.0, 0.0)) == 1
    f.remove_object_location(4, 0.0, 0.0)
    assert len(f.get_objects_unbuffered(0.0, 0.0)) == 0


def test_field_2d_bags():
    """tests/engine/field_2d.rs:186-210"""
    f = ob.Field2D(10.0, 10.0, DISC, TOROIDAL)
    dw, dh, _, _ = f.dims()
    assert (dw, dh) == (21, 21)
    assert len(f.get_empty_bags()) == dh * dw == 441
    f.set_object_location(1, 0.0, 0.0)
    f.set_object_location(2, 0.0, 0.0)
    f.set_object_location(3, 4.0, 4.0)
    f.lazy_update()
    bags = f.get_empty_bags()
    assert len(bags) == dh * dw - 2
    for bx, by in bags:  # get_random_empty_bag can never return an occupied cell origin
        assert (bx, by) != (0.0, 0.0) and (bx, by) != (4.0, 4.0)


def test_field_2d_iter():
    """tests/engine/field_2d.rs:218-247"""
    f = ob.Field2D(10.0, 10.0, DISC, TOROIDAL)
    f.set_object_location(1, 0.0, 0.0)
    f.set_object_location(2, 0.01, 0.01)
    f.set_object_location(3, 5.0, 5.0)
    it = f.iter_objects(unbuffered=True)
    for ox, oy, i in zip(it["ox"], it["oy"], it["id"]):
        assert i in f.get_objects_unbuffered(float(ox), float(oy))
    f.lazy_update()
    assert len(f.get_objects(0.0, 0.0)) == 2
    it = f.iter_objects()
    assert len(it["id"]) == 3
    for ox, oy, i in zip(it["ox"], it["oy"], it["id"]):
        assert i in f.get_objects(float(ox), float(oy))


def test_field_2d_out_of_world_panics():
    """field_2d.rs:838-841: no bounds check -> Vec index panic"""
    f = ob.Field2D(10.0, 10.0, DISC, TOROIDAL)
    with pytest.raises(ob.OraclePanic):
        f.set_object_location(1, -1.0, 0.0)
    with pytest.raises(ob.OraclePanic):
        f.set_object_location(1, 100.0, 100.0)


def test_window_is_clamped_not_wrapped_when_toroidal():
    """field_2d.rs:495-500: toroidal => clamp; non-toroidal => wrap via t_transform"""
    tor = ob.Field2D(10.0, 10.0, 1.0, True)
    non = ob.Field2D(10.0, 10.0, 1.0, False)
    for f in (tor, non):
        f.set_object_location(7, 9.5, 9.5)
        f.lazy_update()
    assert len(tor.get_neighbors_within_relax_distance(0.5, 0.5, 1.0)) == 0
    assert list(non.get_neighbors_within_relax_distance(0.5, 0.5, 1.0)) == [7]
    # exact query on the non-toroidal field filters the wrapped candidate by Euclidean distance
    assert len(non.get_neighbors_within_distance(0.5, 0.5, 1.0)) == 0


def test_padding_cell_is_never_scanned():
    """field_2d.rs:487-488 vs :317-318: agents at x == w sit in the padding column"""
    f = ob.Field2D(10.0, 10.0, 1.0, True)
    f.set_object_location(1, 10.0, 3.0)
    f.lazy_update()
    assert f.num_objects_at_location(10.0, 3.0) == 1
    assert len(f.get_neighbors_within_relax_distance(9.9, 3.0, 5.0)) == 0
    assert len(f.get_neighbors_within_distance(10.0, 3.0, 5.0)) == 0


def test_dense_number_grid_2d_apply():
    """tests/engine/dense_number_grid_2d.rs:31-89"""
    W = H = 10
    g = ob.DenseNumberGrid2D(W, H)
    for i in range(W):
        for j in range(H):
            g.set_value_location(0, i, j)
    g.lazy_update()
    g.apply_const(1, g.WRITE)
    g.lazy_update()
    assert (g.dump() == 1).all()
    g.apply_add(1, g.READWRITE)
    g.lazy_update()
    g.apply_add(1, g.READ)
    for i in range(W):
        for j in range(H):
            assert g.get_value(i, j) == 3
    for i in range(W):
        for j in range(H):
            g.set_value_location(i * j, i, j)
    want = np.outer(np.arange(W), np.arange(H)).astype(np.uint16)
    assert (g.dump(unbuffered=True) == want).all()
    g.lazy_update()
    assert (g.dump() == want).all()


@pytest.mark.parametrize("i,j", [(1, 1), (3, 7), (9, 9), (9, 1)])
def test_dense_number_grid_2d_bags(i, j):
    """tests/engine/dense_number_grid_2d.rs:97-165 ((i,j) = rng.random_range(1..W), (1..H))"""
    W = H = 10
    g = ob.DenseNumberGrid2D(W, H)
    assert g.num_empty_bags() == W * H
    loc = (4, 2)  # stands in for get_random_empty_bag()
    g.set_value_location(10, *loc)
    assert g.get_value_unbuffered(*loc) == 10
    g.remove_value_location(*loc)
    assert g.get_value_unbuffered(*loc) is None
    g.set_value_location(10, *loc)
    g.update()
    assert g.num_empty_bags() == W * H - 1
    for a in range(W):
        for b in range(H):
            g.set_value_location(0, a, b)
    assert g.get_location_unbuffered(0) == (0, 0)
    g.set_value_location(5, i, j)
    assert g.get_location_unbuffered(5) == (i, j)
    assert g.get_location_unbuffered(6) is None
    g.lazy_update()
    assert g.get_location(0) == (0, 0)
    assert g.get_location(5) == (i, j)
    assert g.get_location(6) is None
    assert g.num_empty_bags() == 0


def test_dense_number_grid_lazy_update_clears_unwritten_cells():
    """dense_number_grid_2d.rs:537-545: after the swap every write cell is None"""
    g = ob.DenseNumberGrid2D(4, 3)
    g.set_value_location(7, 1, 2)
    g.lazy_update()
    assert g.get_value(1, 2) == 7
    g.lazy_update()  # nothing written this step
    assert g.get_value(1, 2) is None
    assert g.num_empty_bags() == 12


def test_schedule_operations():
    """tests/engine/schedule.rs:8-39 and :48-85"""
    s = ob.Schedule()
    id0, ok0 = s.schedule_repeating(tag=100)
    id1, ok1 = s.schedule_repeating(tag=101)
    assert (id0, id1) == (0, 1) and ok0 and ok1
    assert list(s.get_all_events()) == [100, 101]  # iteration order = insertion order
    assert s.dequeue(id0)
    assert list(s.get_all_events()) == [101]
    assert not s.dequeue(id0)
    assert s.dequeue(id1)
    assert len(s.get_all_events()) == 0


def test_schedule_pop_order_alternates_for_equal_priorities():
    """schedule.rs:377-407 over priority-queue 2.0.2's swap-remove heap: with all priorities
    equal the pop order is 0,N-1,..,1 on even steps and 0,1,..,N-1 on odd steps."""
    m = fixture_model(6)
    m.init()
    assert list(m.pop_order()) == [0, 5, 4, 3, 2, 1]
    m.step(1)
    assert list(m.pop_order()) == [0, 1, 2, 3, 4, 5]
    m.step(1)
    assert list(m.pop_order()) == [0, 5, 4, 3, 2, 1]


def test_simulate():
    """tests/explore/simulate.rs:18-38: Flockers 200x200, 100 agents, 10 steps run to completion"""
    m = ob.Flockers(200.0, 200.0, 100, DISC, TOROIDAL, ob.boids_params(radius=10.0, exact=1))
    m.init()
    sec = m.time_steps(10)
    assert m.schedule_step == 10 and sec > 0.0
    x, y, dx, dy = m.agents()
    assert np.isfinite(x).all() and (x >= 0).all() and (x <= 200.0).all()
    assert (y >= 0).all() and (y <= 200.0).all()
    # every agent moved by exactly JUMP (bird.rs:139-143) unless it had no displacement at all
    norm = np.sqrt(dx.astype(np.float64) ** 2 + dy.astype(np.float64) ** 2)
    assert np.allclose(norm[norm > 0], 0.7, rtol=1e-5)


def test_flockers_agents_match_field_copies():
    """bird.rs:145-153: the agent's own pos/last_d equal the copy it pushed into the field"""
    m = ob.Flockers(400.0, 400.0, 2000, 10.0 / 1.5, True, ob.boids_params(exact=0))
    m.init()
    m.step(3)
    x, y, dx, dy = m.agents()
    it = m.field1.iter_objects()
    assert sorted(it["id"]) == list(range(2000))
    assert (x[it["id"]] == it["x"]).all() and (y[it["id"]] == it["y"]).all()
    assert (dx[it["id"]] == it["ldx"]).all() and (dy[it["id"]] == it["ldy"]).all()
