"""GPU `SparseGrid2D` — the hash-map object grid of src/engine/fields/sparse_object_grid_2d.rs:203-721 on the
device (csrc/objgrid.cu, sparse mode: bags are slots of a device key table).  Same method names and argument
meaning as the reference; objects are `(id, tag)` pairs that compare by id.  Any `(x, y)` is a key; iteration
order is unspecified, as with the reference's HashMap."""
import ctypes as C

import numpy as np

from ... import _abi as abi
from .dense_object_grid_2d import DenseGrid2D


class SparseGrid2D(DenseGrid2D):
    def __init__(self, width, height, capacity=1 << 16, device=0):
        """SparseGrid2D::new(width, height)  :224-234 (+ device capacity in objects)"""
        self._h = abi.vp()
        abi.check(abi.lib().kg_objgrid_create_sparse(width, height, capacity, device, C.byref(self._h)))
        self.width, self.height = width, height

    def set_object_location(self, object, loc):
        """:648-659 — pushed; an equal object already in the bag stays"""
        super().set_object_location(object, loc)

    def update(self):
        """Field::update :711-718 — the read map becomes a copy of the write map, the write map is cleared"""
        abi.check(abi.lib().kg_objgrid_update(self._h))

    def bag_sizes(self, unbuffered=False):
        """objects per cell of the nominal [0, width) x [0, height) area"""
        w, h = max(self.width, 0), max(self.height, 0)
        out = np.zeros(max(w * h, 1), np.uint32)
        abi.check(abi.lib().kg_objgrid_bag_sizes(self._h, int(unbuffered), len(out), abi.ptr(out)))
        return out[:w * h].reshape(w, h)

    def get_empty_bags(self):
        """:482-499 — cells of the area without a key (or with an empty bag), x outer, y inner"""
        xs, ys = np.nonzero(self.bag_sizes() == 0)
        return [(int(a), int(b)) for a, b in zip(xs, ys)]
