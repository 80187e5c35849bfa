"""The reference's known-answer tests for DenseGrid2D (tests/engine/dense_object_grid_2d.rs:31-180)
and SparseGrid2D (tests/engine/sparse_object_grid_2d.rs) and the doc examples of
src/engine/fields/{dense,sparse}_object_grid_2d.rs, run on the oracle's restatement
(oracle/object_grid.hpp).  SURVEY §8(f) rank 2: the oracle for the next field comes before its
device side.  Objects are (id, tag); equality is by id, as for the fixture's Bird."""
import random

import pytest

import oracle_binding as ob

WIDTH = HEIGHT = 10
G = ob.DenseGrid2D


def test_dense_object_grid_2d_bags():
    """tests/engine/dense_object_grid_2d.rs:31-104"""
    grid = G(WIDTH, HEIGHT)
    assert len(grid.get_empty_bags()) == 100
    loc = random.Random(1).choice(grid.get_empty_bags())      # get_random_empty_bag (:391-400)
    grid.set_object_location((0, 0), loc)
    assert grid.get_location_unbuffered((0, 0)) == loc
    assert grid.get_location_unbuffered((1, 0)) is None
    assert grid.get_location((0, 0)) is None                    # not visible before the swap
    grid.lazy_update()
    assert grid.get_location((0, 0)) == loc
    assert grid.get_location((1, 0)) is None
    assert len(grid.get_empty_bags()) == 99
    for i in range(HEIGHT):
        for j in range(WIDTH):
            grid.set_object_location((i * HEIGHT + j, 0), (i, j))
    grid.lazy_update()
    assert len(grid.get_empty_bags()) == 0


def test_dense_object_grid_2d_apply():
    """tests/engine/dense_object_grid_2d.rs:112-180"""
    grid = G(WIDTH, HEIGHT)
    for i in range(HEIGHT):
        for j in range(WIDTH):
            grid.set_object_location((i * HEIGHT + j, 0), (i, j))
    seen = grid.iter_objects_unbuffered()
    assert len(seen) == 100
    for loc, (oid, _) in seen:
        assert grid.get_objects_unbuffered(loc)[0][0] == oid
    assert grid.iter_objects() == []
    grid.lazy_update()
    seen = grid.iter_objects()
    assert len(seen) == 100
    for loc, (oid, _) in seen:
        assert grid.get_objects(loc)[0][0] == oid
    # flag = true through WRITE: results land in the write bags
    assert grid.apply_to_all_values(G.SET_TAG, 1, G.WRITE) == 100
    assert all(tag == 1 for _, (_, tag) in grid.iter_objects_unbuffered())
    assert len(grid.iter_objects_unbuffered()) == 100
    # flag = false through READ: the read bags are rewritten in place
    grid.apply_to_all_values(G.SET_TAG, 0, G.READ)
    assert all(tag == 0 for _, (_, tag) in grid.iter_objects())
    # READWRITE: the write bags are not empty, so their elements are updated in place
    grid.apply_to_all_values(G.SET_TAG, 1, G.READWRITE)
    grid.lazy_update()
    assert len(grid.iter_objects()) == 100
    assert all(tag == 1 for _, (_, tag) in grid.iter_objects())


def test_doc_example_apply_none_removes():
    """:233-257: WRITE copies the read objects into the write bags; returning None from a READ
    pass deletes the objects"""
    grid = G(10, 10)
    for i in range(10):
        for j in range(10):
            grid.set_object_location((i * 10 + j, 0), (i, j))
    grid.lazy_update()
    grid.apply_to_all_values(G.SET_TAG, 1, G.WRITE)
    grid.lazy_update()
    assert all(tag == 1 for _, (_, tag) in grid.iter_objects())
    grid.apply_to_all_values(G.REMOVE, 0, G.READ)
    assert grid.iter_objects() == [] and len(grid.get_empty_bags()) == 100


def test_set_object_location_replaces_an_equal_object():
    """:688-697: `retain(obj != object)` then push — one copy per bag, moved to the back, with
    the new payload; a different bag keeps its own copy"""
    grid = G(4, 4)
    grid.set_object_location((7, 1), (2, 3))
    grid.set_object_location((8, 0), (2, 3))
    grid.set_object_location((7, 5), (2, 3))
    assert grid.get_objects_unbuffered((2, 3)) == [(8, 0), (7, 5)]
    grid.set_object_location((7, 9), (0, 0))
    assert grid.get_location_unbuffered((7, 0)) == (0, 0)     # first bag in x-major order
    grid.remove_object_location((7, 0), (2, 3))               # :729-736
    assert grid.get_objects_unbuffered((2, 3)) == [(8, 0)]
    grid.remove_object_location((8, 0), (2, 3))
    assert grid.get_objects_unbuffered((2, 3)) is None          # Option::None for an empty bag
    assert grid.get_objects((2, 3)) is None


def test_lazy_update_swaps_and_clears_the_write_bags():
    """:743-750"""
    grid = G(3, 5)
    grid.set_object_location((1, 0), (2, 4))
    grid.lazy_update()
    assert grid.get_objects((2, 4)) == [(1, 0)] and grid.iter_objects_unbuffered() == []
    grid.lazy_update()                                          # nothing was written: all gone
    assert grid.iter_objects() == [] and grid.iter_objects_unbuffered() == []


def test_iteration_order_is_x_outer_y_inner_then_bag_order():
    """:589-608"""
    grid = G(3, 2)
    for oid, loc in [(5, (2, 1)), (1, (0, 1)), (9, (0, 0)), (2, (0, 1)), (4, (1, 0))]:
        grid.set_object_location((oid, 0), loc)
    grid.lazy_update()
    assert [(loc, oid) for loc, (oid, _) in grid.iter_objects()] == \
        [((0, 0), 9), ((0, 1), 1), ((0, 1), 2), ((1, 0), 4), ((2, 1), 5)]
    assert grid.get_empty_bags() == [(1, 1), (2, 0)]


def test_readwrite_takes_read_objects_only_where_the_write_bag_is_empty():
    """:293-326: a non-empty write bag is updated in place and the read bag is ignored; an empty
    one receives the closure's results for the read bag, without duplicates"""
    grid = G(2, 2)
    grid.set_object_location((1, 0), (0, 0))
    grid.set_object_location((2, 0), (1, 1))
    grid.lazy_update()
    grid.set_object_location((3, 0), (0, 0))                  # write bag (0,0) now holds only 3
    calls = grid.apply_to_all_values(G.SET_TAG, 7, G.READWRITE)
    assert calls == 2                                           # object 3 (write) and object 2 (read)
    assert grid.get_objects_unbuffered((0, 0)) == [(3, 7)]
    assert grid.get_objects_unbuffered((1, 1)) == [(2, 7)]
    assert grid.get_objects((0, 0)) == [(1, 0)]               # the read side is untouched


def test_quirk_closure_bag_id_is_y_major():
    """:768-779: apply_to_all_values hands the closure calculate_indexes_bag(i) = (i - width*row,
    row); every other method reads flat index i as (i / height, i % height).  They agree on the
    diagonal of a square grid only."""
    grid = G(3, 2)                                              # width 3, height 2
    for x in range(3):
        for y in range(2):
            grid.set_object_location((x * 2 + y, 0), (x, y))
    grid.lazy_update()
    grid.apply_to_all_values(G.TAG_WITH_BAG_ID, 0, G.READ)
    got = {loc: (tag >> 16, tag & 0xFFFF) for loc, (_, tag) in grid.iter_objects()}
    for (x, y), bag_id in got.items():
        i = x * 2 + y
        assert bag_id == (i - 3 * (i // 3), i // 3)
    assert got[(0, 1)] == (1, 0) and got[(1, 0)] == (2, 0) and got[(2, 1)] == (2, 1)


def test_quirk_update_inserts_instead_of_assigning():
    """:753-763: Vec::insert grows the read Vec — the first width*height bags become copies of the
    write bags (so reads see the unswapped writes), the old read bags are pushed behind them, and
    apply_to_all_values then panics on the first index it cannot map to a bag"""
    grid = G(3, 3)
    grid.set_object_location((1, 0), (0, 0))
    grid.lazy_update()
    grid.set_object_location((2, 0), (1, 2))
    grid.update()
    assert grid.nbags() == 18 and grid.nbags(unbuffered=True) == 9
    assert grid.get_objects((1, 2)) == [(2, 0)] and grid.get_objects((0, 0)) is None
    assert grid.get_objects_unbuffered((1, 2)) == [(2, 0)]     # update() does not clear the writes
    with pytest.raises(ob.OraclePanic):
        grid.apply_to_all_values(G.SET_TAG, 1, G.READ)


def test_out_of_grid_location_panics():
    grid = G(4, 4)
    with pytest.raises(ob.OraclePanic):
        grid.set_object_location((1, 0), (4, 0))
    with pytest.raises(ob.OraclePanic):
        grid.get_objects((0, 16))
    grid.set_object_location((1, 0), (0, 5))                  # flat index 5 = cell (1, 1): no bounds check per axis
    assert grid.get_location_unbuffered((1, 0)) == (1, 1)


# ---------------------------------------------------------------- SparseGrid2D
S = ob.SparseGrid2D


def test_sparse_object_grid_2d_bags():
    """tests/engine/sparse_object_grid_2d.rs: sparse_object_grid_2d_bags"""
    grid = S(WIDTH, HEIGHT)
    assert len(grid.get_empty_bags()) == 100
    loc = random.Random(2).choice(grid.get_empty_bags())      # stands in for get_random_empty_bag
    grid.set_object_location((0, 0), loc)
    assert grid.get_location_unbuffered((0, 0)) == loc
    assert grid.get_location_unbuffered((3, 0)) is None
    grid.update()                                               # copy write -> read, clear write
    assert grid.get_location((0, 0)) == loc
    assert grid.get_location((3, 0)) is None
    assert grid.get_location_unbuffered((0, 0)) is None
    assert len(grid.get_empty_bags()) == 99
    for i in range(HEIGHT):
        for j in range(WIDTH):
            grid.set_object_location((i * HEIGHT + j, 0), (i, j))
    grid.lazy_update()
    assert len(grid.get_empty_bags()) == 0
    grid.lazy_update()                                          # clear all objects
    loc = random.Random(3).choice(grid.get_empty_bags())
    grid.set_object_location((0, 0), loc)
    grid.remove_object_location((0, 0), loc)
    assert grid.get_objects_unbuffered(loc) is None             # the emptied bag lost its key
    grid.lazy_update()
    assert len(grid.get_empty_bags()) == 100


def test_sparse_object_grid_2d_apply():
    """tests/engine/sparse_object_grid_2d.rs: sparse_object_grid_2d_apply"""
    grid = S(WIDTH, HEIGHT)
    for i in range(HEIGHT):
        for j in range(WIDTH):
            grid.set_object_location((i * HEIGHT + j, 0), (i, j))
    seen = grid.iter_objects_unbuffered()
    assert len(seen) == 100
    for loc, (oid, _) in seen:
        assert grid.get_objects_unbuffered(loc)[0][0] == oid
    grid.lazy_update()
    seen = grid.iter_objects()
    assert sorted(oid for _, (oid, _) in seen) == list(range(100))
    for loc, (oid, _) in seen:
        assert grid.get_objects(loc)[0][0] == oid
    # WRITE walks the write map, which is empty after the swap: no closure call
    assert grid.apply_to_all_values(S.SET_TAG, 1, S.WRITE) == 0
    assert grid.iter_objects_unbuffered() == []
    grid.apply_to_all_values(S.SET_TAG, 0, S.READ)
    assert all(tag == 0 for _, (_, tag) in grid.iter_objects())
    # READWRITE: no write bag for any read key -> one new bag per key with the closure's result
    assert grid.apply_to_all_values(S.SET_TAG, 1, S.READWRITE) == 100
    grid.lazy_update()
    assert len(grid.iter_objects()) == 100 and all(tag == 1 for _, (_, tag) in grid.iter_objects())


def test_sparse_differences_from_the_dense_grid():
    """:648-659 push without replacing; any Int2D is a key; :278-320 a closure returning None
    panics; READWRITE keeps one object per NEW write bag"""
    grid = S(4, 4)
    grid.set_object_location((7, 1), (2, 3))
    grid.set_object_location((7, 2), (2, 3))
    assert grid.get_objects_unbuffered((2, 3)) == [(7, 1), (7, 2)]
    grid.set_object_location((9, 0), (-5, 1000))               # outside [0,4)^2: still stored
    assert grid.get_location_unbuffered((9, 0)) == (-5, 1000)
    grid.lazy_update()
    assert grid.get_empty_bags() == [(x, y) for x in range(4) for y in range(4) if (x, y) != (2, 3)]
    with pytest.raises(ob.OraclePanic):
        grid.apply_to_all_values(S.REMOVE, 0, S.READ)
    grid.apply_to_all_values(S.SET_TAG, 5, S.READWRITE)
    assert grid.get_objects_unbuffered((2, 3)) == [(7, 5)]      # the last of the two equal-key inserts
    assert grid.get_objects((2, 3)) == [(7, 1), (7, 2)]
