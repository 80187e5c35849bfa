"""GPU `DenseGrid2D` — the object grid of src/engine/fields/dense_object_grid_2d.rs:175-779 (default
variant) on the device's cell-sorted layout (csrc/objgrid.cu).  Same method names and argument meaning
as the reference; objects are `(id, tag)` pairs that compare by id, like the fixture's Bird (the tag
stands in for `Bird.flag`).  Failures raise where the reference panics."""
import ctypes as C

import random

import numpy as np

from ... import _abi as abi
from .grid_option import GridOption
from .field import Field


class DenseGrid2D(Field):
    READ, WRITE, READWRITE = GridOption.READ, GridOption.WRITE, GridOption.READWRITE
    # the closure family apply_to_all_values accepts (include/krabgpu.h KG_OBJ_*)
    SET_TAG, REMOVE, REMOVE_IF_TAG, TAG_WITH_BAG_ID = 0, 1, 2, 3

    def __init__(self, width, height, capacity=1 << 16, device=0):
        """DenseGrid2D::new(width, height)  :201-214 (+ device capacity in objects)"""
        self._h = abi.vp()
        abi.check(abi.lib().kg_objgrid_create(width, height, capacity, device, C.byref(self._h)))
        self.width, self.height = abs(width), abs(height)

    def close(self):
        if getattr(self, "_h", None):
            abi.lib().kg_objgrid_destroy(self._h)
            self._h = None

    __del__ = close

    # ------------------------------------------------------------------ write side
    def set_object_location(self, object, loc):
        """:688-697 — an equal object already in that write bag is replaced"""
        self.set_object_locations([object[0]], [object[1] if len(object) > 1 else 0], [loc[0]], [loc[1]])

    def set_object_locations(self, ids, tags, xs, ys):
        ids, tags = abi.as_u32(ids), abi.as_u32(tags)
        xs, ys = abi.as_i32(xs), abi.as_i32(ys)
        assert len(ids) == len(tags) == len(xs) == len(ys)
        abi.check(abi.lib().kg_objgrid_set_object_locations(self._h, len(ids), abi.ptr(ids), abi.ptr(tags),
                                                            abi.ptr(xs), abi.ptr(ys)))

    def remove_object_location(self, object, loc):
        """:729-736"""
        ids, xs, ys = abi.as_u32([object[0]]), abi.as_i32([loc[0]]), abi.as_i32([loc[1]])
        abi.check(abi.lib().kg_objgrid_remove_object_locations(self._h, 1, abi.ptr(ids), abi.ptr(xs), abi.ptr(ys)))

    def lazy_update(self):
        """Field::lazy_update :743-750"""
        abi.check(abi.lib().kg_objgrid_lazy_update(self._h))

    def update(self):
        """Field::update :753-763 — refused by the device (see include/krabgpu.h)"""
        abi.check(abi.lib().kg_objgrid_update(self._h))

    # ------------------------------------------------------------------ read side
    def num_objects(self, unbuffered=False):
        out = abi.u64()
        abi.check(abi.lib().kg_objgrid_num_objects(self._h, int(unbuffered), C.byref(out)))
        return out.value

    def get_objects(self, loc, unbuffered=False):
        """:507-520 — list of (id, tag) in bag order, None for an empty bag"""
        cap = 64
        while True:
            ids, tags, n = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32), abi.u64()
            rc = abi.lib().kg_objgrid_get_objects(self._h, int(unbuffered), loc[0], loc[1], cap, abi.ptr(ids),
                                                  abi.ptr(tags), C.byref(n))
            if rc == abi.KG_E_CAPACITY and n.value > cap:
                cap = n.value
                continue
            abi.check(rc)
            if n.value == 0:
                return None
            return [(int(a), int(b)) for a, b in zip(ids[:n.value], tags[:n.value])]

    def get_objects_unbuffered(self, loc):
        """:547-561"""
        return self.get_objects(loc, True)

    def get_location(self, object, unbuffered=False):
        """:429-441 — first bag (x outer, y inner) holding an equal object"""
        x, y, found = abi.i32(), abi.i32(), C.c_int()
        abi.check(abi.lib().kg_objgrid_get_location(self._h, int(unbuffered), int(object[0]), C.byref(x), C.byref(y),
                                                    C.byref(found)))
        return (x.value, y.value) if found.value else None

    def get_location_unbuffered(self, object):
        """:471-482"""
        return self.get_location(object, True)

    def bag_sizes(self, unbuffered=False):
        n = self.width * self.height
        out = np.zeros(max(n, 1), np.uint32)
        abi.check(abi.lib().kg_objgrid_bag_sizes(self._h, int(unbuffered), len(out), abi.ptr(out)))
        return out[:n].reshape(self.width, self.height)

    def get_empty_bags(self):
        """:358-370"""
        xs, ys = np.nonzero(self.bag_sizes() == 0)
        return [(int(a), int(b)) for a, b in zip(xs, ys)]

    def get_random_empty_bag(self, rng=None):
        """:391-400"""
        bags = self.get_empty_bags()
        return bags[(rng or random).randrange(len(bags))] if bags else None

    def iter_objects(self, unbuffered=False):
        """:589-608 — [((x, y), (id, tag)), ...] in the order the reference calls the closure"""
        n = self.num_objects(unbuffered)
        xs, ys = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
        ids, tags, got = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.uint32), abi.u64()
        abi.check(abi.lib().kg_objgrid_iter_objects(self._h, int(unbuffered), len(xs), abi.ptr(xs), abi.ptr(ys),
                                                    abi.ptr(ids), abi.ptr(tags), C.byref(got)))
        return [((int(xs[i]), int(ys[i])), (int(ids[i]), int(tags[i]))) for i in range(got.value)]

    def iter_objects_unbuffered(self):
        """:634-654"""
        return self.iter_objects(True)

    def apply_to_all_values(self, closure, arg, option):
        """:258-328 for closure in {SET_TAG, REMOVE, REMOVE_IF_TAG, TAG_WITH_BAG_ID}; returns the
        number of closure calls"""
        calls = abi.u64()
        abi.check(abi.lib().kg_objgrid_apply(self._h, int(closure), int(arg), int(GridOption(option)),
                                             C.byref(calls)))
        return calls.value
