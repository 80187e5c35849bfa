"""DenseGrid2D on the device (csrc/objgrid.cu) against the oracle's restatement of
src/engine/fields/dense_object_grid_2d.rs: the reference's own known-answer tests
(tests/engine/dense_object_grid_2d.rs:31-180, ported in test_oracle_object_grid.py) run unchanged on
the GPU class, and random op sequences are compared with the oracle after every call."""
import random

import numpy as np
import pytest

import krabmaga_b200 as kb
import oracle_binding as ob
import test_oracle_object_grid as kats
from krabmaga_b200 import _abi as abi

pytestmark = pytest.mark.gpu

KAT_NAMES = ["test_dense_object_grid_2d_bags", "test_dense_object_grid_2d_apply",
             "test_doc_example_apply_none_removes", "test_set_object_location_replaces_an_equal_object",
             "test_lazy_update_swaps_and_clears_the_write_bags",
             "test_iteration_order_is_x_outer_y_inner_then_bag_order",
             "test_readwrite_takes_read_objects_only_where_the_write_bag_is_empty",
             "test_quirk_closure_bag_id_is_y_major"]


@pytest.mark.parametrize("name", KAT_NAMES)
def test_reference_kats_on_the_device(name, monkeypatch):
    """the very test bodies that pin the oracle, with G = the GPU DenseGrid2D"""
    monkeypatch.setattr(kats, "G", kb.DenseGrid2D)
    getattr(kats, name)()


def test_out_of_grid_and_update():
    g = kb.DenseGrid2D(4, 4)
    with pytest.raises(kb.KgOutOfBounds):
        g.set_object_location((1, 0), (4, 0))
    with pytest.raises(kb.KgOutOfBounds):
        g.get_objects((0, 16))
    g.set_object_location((1, 0), (0, 5))        # flat index 5 = bag (1, 1): no per-axis check, as in the reference
    assert g.get_location_unbuffered((1, 0)) == (1, 1)
    with pytest.raises(kb.KgError) as e:
        g.update()                                # the reference's update() corrupts the read Vec: refused
    assert e.value.code == abi.KG_E_INVALID
    with pytest.raises(kb.KgError):
        kb.DenseGrid2D(1 << 16, 1 << 16)
    g.close()


def test_capacity_is_enforced_and_replaced_objects_free_their_slots():
    g = kb.DenseGrid2D(4, 4, capacity=8)
    for k in range(40):                           # 40 inserts of 4 distinct objects: replace-on-insert
        g.set_object_location((k % 4, k), (1, 2))
    assert g.get_objects_unbuffered((1, 2)) == [(0, 36), (1, 37), (2, 38), (3, 39)]
    with pytest.raises(kb.KgError) as e:
        g.set_object_locations(list(range(10, 20)), [0] * 10, [0] * 10, [0] * 10)
    assert e.value.code == abi.KG_E_CAPACITY
    g.close()


@pytest.mark.parametrize("seed,w,h", [(1, 5, 5), (2, 7, 3), (3, 2, 9), (4, 12, 12)])
def test_random_op_sequences_match_the_oracle(seed, w, h):
    rng = random.Random(seed)
    dev, ora = kb.DenseGrid2D(w, h, capacity=4096), ob.DenseGrid2D(w, h)
    G = ob.DenseGrid2D
    for step in range(160):
        r = rng.random()
        if r < 0.55:
            n = rng.randint(1, 6)
            for _ in range(n):
                obj, loc = (rng.randint(0, 14), rng.randint(0, 9)), (rng.randrange(w), rng.randrange(h))
                dev.set_object_location(obj, loc)
                ora.set_object_location(obj, loc)
        elif r < 0.70:
            obj, loc = (rng.randint(0, 14), 0), (rng.randrange(w), rng.randrange(h))
            dev.remove_object_location(obj, loc)
            ora.remove_object_location(obj, loc)
        elif r < 0.88:
            op = rng.choice([G.SET_TAG, G.REMOVE, G.REMOVE_IF_TAG, G.TAG_WITH_BAG_ID])
            option = rng.choice([G.READ, G.WRITE, G.READWRITE])
            arg = rng.randint(0, 9)
            assert dev.apply_to_all_values(op, arg, option) == ora.apply_to_all_values(op, arg, option), step
        else:
            dev.lazy_update()
            ora.lazy_update()
        if step % 3 == 0 or r >= 0.70:
            assert dev.iter_objects() == ora.iter_objects(), step
            assert dev.iter_objects_unbuffered() == ora.iter_objects_unbuffered(), step
            assert dev.get_empty_bags() == ora.get_empty_bags(), step
            probe = (rng.randint(0, 14), 0)
            assert dev.get_location(probe) == ora.get_location(probe), step
            assert dev.get_location_unbuffered(probe) == ora.get_location_unbuffered(probe), step
            loc = (rng.randrange(w), rng.randrange(h))
            assert dev.get_objects(loc) == ora.get_objects(loc) and \
                dev.get_objects_unbuffered(loc) == ora.get_objects_unbuffered(loc), step
    dev.close()


def test_a_schelling_sized_grid_round_trips():
    """200 x 200 bags, 30,000 objects in one call, moved once: counts, locations and order"""
    w = h = 200
    rng = np.random.default_rng(5)
    n = 30000
    ids = np.arange(n, dtype=np.uint32)
    xs, ys = rng.integers(0, w, n).astype(np.int32), rng.integers(0, h, n).astype(np.int32)
    g = kb.DenseGrid2D(w, h, capacity=4 * n)
    g.set_object_locations(ids, ids % 3, xs, ys)
    g.lazy_update()
    sizes = g.bag_sizes()
    want = np.zeros((w, h), np.int64)
    np.add.at(want, (xs, ys), 1)
    assert (sizes == want).all() and g.num_objects() == n
    it = g.iter_objects()
    cells = [loc[0] * h + loc[1] for loc, _ in it]
    assert cells == sorted(cells)                                         # x outer, y inner
    order = np.lexsort((np.arange(n), xs.astype(np.int64) * h + ys))      # bag order = insertion order
    assert [o[0] for _, o in it] == list(ids[order])
    assert g.get_location((123, 0)) == (int(xs[123]), int(ys[123]))
    assert g.apply_to_all_values(kb.DenseGrid2D.REMOVE_IF_TAG, 0, kb.DenseGrid2D.READ) == n
    assert g.num_objects() == int((ids % 3 != 0).sum())
    g.close()


# ---------------------------------------------------------------- SparseGrid2D on the device
SPARSE_KATS = ["test_sparse_object_grid_2d_bags", "test_sparse_object_grid_2d_apply",
               "test_sparse_differences_from_the_dense_grid"]


@pytest.mark.parametrize("name", SPARSE_KATS)
def test_sparse_reference_kats_on_the_device(name, monkeypatch):
    """tests/engine/sparse_object_grid_2d.rs as ported for the oracle, with S = the GPU SparseGrid2D; the
    oracle's panic is the device's KgError"""
    monkeypatch.setattr(kats, "S", kb.SparseGrid2D)
    monkeypatch.setattr(ob, "OraclePanic", kb.KgError)
    getattr(kats, name)()


def canon(it):
    """iteration order of a HashMap is unspecified: bags compared as key -> object list"""
    out = {}
    for loc, obj in it:
        out.setdefault(loc, []).append(obj)
    return out


@pytest.mark.parametrize("seed,w,h,span", [(1, 5, 5, 5), (2, 7, 3, 40), (3, 4, 9, 100000)])
def test_sparse_random_op_sequences_match_the_oracle(seed, w, h, span):
    """random calls (keys inside and far outside width x height, negative ones included) compared with the
    oracle after every few calls: every bag, in bag order"""
    rng = random.Random(seed)
    dev, ora = kb.SparseGrid2D(w, h, capacity=64), ob.SparseGrid2D(w, h)   # small capacity: the key table is rebuilt often
    S = ob.SparseGrid2D
    keys = [(rng.randint(-span, span), rng.randint(-span, span)) for _ in range(12)] + \
           [(rng.randrange(w), rng.randrange(h)) for _ in range(12)]
    for step in range(400):
        r = rng.random()
        if r < 0.5:
            for _ in range(rng.randint(1, 3)):
                obj, loc = (rng.randint(0, 9), rng.randint(0, 9)), rng.choice(keys)
                if rng.random() < 0.6:          # a key never seen before: the table fills up and is rebuilt
                    loc = (rng.randint(-span, span), rng.randint(-span, span))
                    keys[rng.randrange(len(keys))] = loc
                if sum(len(v) for v in canon(ora.iter_objects_unbuffered()).values()) < 40:
                    dev.set_object_location(obj, loc)
                    ora.set_object_location(obj, loc)
        elif r < 0.72:
            obj, loc = (rng.randint(0, 9), 0), rng.choice(keys)
            dev.remove_object_location(obj, loc)
            ora.remove_object_location(obj, loc)
        elif r < 0.86:
            op = rng.choice([S.SET_TAG, S.TAG_WITH_BAG_ID, S.REMOVE_IF_TAG])
            option = rng.choice([S.READ, S.WRITE, S.READWRITE])
            arg = rng.randint(20, 29) if op == S.REMOVE_IF_TAG else rng.randint(0, 9)   # tag never present: Some(..)
            assert dev.apply_to_all_values(op, arg, option) == ora.apply_to_all_values(op, arg, option), step
        elif r < 0.95:
            dev.lazy_update()
            ora.lazy_update()
        else:
            dev.update()
            ora.update()
        assert canon(dev.iter_objects()) == canon(ora.iter_objects()), step
        assert canon(dev.iter_objects_unbuffered()) == canon(ora.iter_objects_unbuffered()), step
        if step % 4 == 0:
            assert dev.get_empty_bags() == ora.get_empty_bags(), step
            loc = rng.choice(keys)
            assert dev.get_objects(loc) == ora.get_objects(loc), step
            assert dev.get_objects_unbuffered(loc) == ora.get_objects_unbuffered(loc), step
            probe = (rng.randint(0, 9), 0)
            for unb in (False, True):
                where = dev.get_location(probe, unb)
                holders = [loc for loc, objs in canon(ora.iter_objects(unb)).items() if any(o[0] == probe[0] for o in objs)]
                assert (where is None and not holders) or where in holders, step
    dev.close()


def test_sparse_closure_none_is_the_reference_panic_and_capacity():
    g = kb.SparseGrid2D(3, 3, capacity=16)
    g.set_object_location((1, 5), (0, 0))
    g.lazy_update()
    with pytest.raises(kb.KgError) as e:
        g.apply_to_all_values(kb.SparseGrid2D.REMOVE_IF_TAG, 5, kb.SparseGrid2D.READ)
    assert e.value.code == abi.KG_E_INVALID
    with pytest.raises(kb.KgError) as e:
        g.set_object_locations(list(range(40)), [0] * 40, [0] * 40, [0] * 40)
    assert e.value.code == abi.KG_E_CAPACITY
    with pytest.raises(kb.KgError):
        g.set_object_location((1, 0), (2**31 - 1, 2**31 - 1))   # the one reserved key
    g.close()


def test_sparse_large_round_trip():
    """100,000 objects on 60,000 distinct keys scattered over +-10^6: every bag, in insertion order"""
    rng = np.random.default_rng(11)
    n = 100000
    kx = rng.integers(-10**6, 10**6, 60000).astype(np.int32)
    ky = rng.integers(-10**6, 10**6, 60000).astype(np.int32)
    pick = rng.integers(0, 60000, n)
    ids = np.arange(n, dtype=np.uint32)
    g = kb.SparseGrid2D(100, 100, capacity=2 * n)
    g.set_object_locations(ids, ids % 7, kx[pick], ky[pick])
    g.lazy_update()
    assert g.num_objects() == n
    got = canon(g.iter_objects())
    want = {}
    for i in range(n):
        want.setdefault((int(kx[pick[i]]), int(ky[pick[i]])), []).append((i, i % 7))
    assert got == want
    g.close()
