"""GPU `DenseNumberGrid2D<T>` — same method names as the reference
(src/engine/fields/dense_number_grid_2d.rs:90-561).  `Option<T>` crosses the boundary as T with one
reserved value (`none`, default T::MAX) meaning None.  Closures cannot cross a C ABI, so
`apply_to_all_values` takes one of the shipped closure families instead of a Rust `Fn(&T)->T`.
"""
import ctypes as C

import random

import numpy as np

from ... import _abi as abi
from ..location import Int2D
from .field import Field
from .grid_option import GridOption

_DTYPES = {1: np.uint8, 2: np.uint16, 4: np.uint32}


class DenseNumberGrid2D(Field):
    def __init__(self, width, height, elem_size=1, none=None, device=0):
        """DenseNumberGrid2D::new(width, height)  :112-126"""
        self.dtype = np.dtype(_DTYPES[elem_size])
        self.none = int(np.iinfo(self.dtype).max) if none is None else int(none)
        self._h = abi.vp()
        abi.check(abi.lib().kg_grid_create(width, height, elem_size, self.none, device, C.byref(self._h)))
        self.width, self.height = abs(width), abs(height)

    def close(self):
        if getattr(self, "_h", None):
            abi.lib().kg_grid_destroy(self._h)
            self._h = None

    __del__ = close

    def _opt(self, v):
        return None if int(v) == self.none else int(v)

    def set_value_location(self, value, loc):
        """:492-495"""
        self.set_values([loc[0]], [loc[1]], [value])

    def set_values(self, xs, ys, values):
        xs, ys = abi.as_i32(xs), abi.as_i32(ys)
        v = np.ascontiguousarray(values, dtype=self.dtype)
        abi.check(abi.lib().kg_grid_set_values(self._h, len(xs), abi.ptr(xs), abi.ptr(ys), abi.ptr(v)))

    def remove_value_location(self, loc):
        """:526-530"""
        xs, ys = abi.as_i32([loc[0]]), abi.as_i32([loc[1]])
        abi.check(abi.lib().kg_grid_remove_values(self._h, 1, abi.ptr(xs), abi.ptr(ys)))

    def get_values(self, xs, ys, unbuffered=False):
        xs, ys = abi.as_i32(xs), abi.as_i32(ys)
        out = np.zeros(len(xs), self.dtype)
        abi.check(abi.lib().kg_grid_get_values(self._h, int(unbuffered), len(xs), abi.ptr(xs),
                                               abi.ptr(ys), abi.ptr(out)))
        return out

    def get_value(self, loc):
        """:350-354"""
        return self._opt(self.get_values([loc[0]], [loc[1]])[0])

    def get_value_unbuffered(self, loc):
        """:376-380"""
        return self._opt(self.get_values([loc[0]], [loc[1]], True)[0])

    def apply_to_all_values(self, closure, option):
        """:155-195.  closure = ("const", c) for |_| c, ("add", c) for |v| v + c, or — any closure — its body
        as a CUDA C expression in `v` (also x, y), e.g. "v - 1", compiled at run time for the device."""
        if isinstance(closure, str):
            abi.check(abi.lib().kg_grid_apply_expr(self._h, closure.encode(), int(GridOption(option))))
            return
        kind, c = closure
        op = {"const": abi.KG_APPLY_CONST, "add": abi.KG_APPLY_ADD}[kind]
        abi.check(abi.lib().kg_grid_apply(self._h, op, int(c), int(GridOption(option))))

    def get_location(self, value, unbuffered=False):
        """:204-229: first match in x-outer / y-inner order."""
        x, y, found = abi.i32(), abi.i32(), C.c_int()
        abi.check(abi.lib().kg_grid_get_location(self._h, int(unbuffered), int(value), C.byref(x),
                                                 C.byref(y), C.byref(found)))
        return Int2D(x.value, y.value) if found.value else None

    def get_location_unbuffered(self, value):
        return self.get_location(value, True)

    def num_empty_bags(self):
        out = abi.u64()
        abi.check(abi.lib().kg_grid_num_empty(self._h, C.byref(out)))
        return out.value

    def get_empty_bags(self):
        """:236-247"""
        g = self.download()
        xs, ys = np.nonzero(g == self.none)
        return [Int2D(int(a), int(b)) for a, b in zip(xs, ys)]

    def get_random_empty_bag(self, rng=None):
        """:319-328: one of get_empty_bags() drawn uniformly, None when the read grid is full."""
        bags = self.get_empty_bags()
        if not bags:
            return None
        return bags[(rng or random).randrange(len(bags))]

    def iter_values(self, closure, unbuffered=False):
        """:404-452: closure(loc: Int2D, value) for every Some cell, x outer / y inner."""
        g = self.download(unbuffered)
        xs, ys = np.nonzero(g != self.none)
        for a, b in zip(xs, ys):
            closure(Int2D(int(a), int(b)), int(g[a, b]))

    def iter_values_unbuffered(self, closure):
        self.iter_values(closure, True)

    def lazy_update(self):
        """:537-545"""
        abi.check(abi.lib().kg_grid_lazy_update(self._h))

    def update(self):
        """:553-561"""
        abi.check(abi.lib().kg_grid_update(self._h))

    def upload(self, cells, unbuffered=False):
        c = np.ascontiguousarray(cells, dtype=self.dtype).reshape(-1)
        assert c.size == self.width * self.height
        abi.check(abi.lib().kg_grid_upload(self._h, int(unbuffered), abi.ptr(c)))

    def download(self, unbuffered=False, out=None):
        if out is None:
            out = np.zeros(self.width * self.height, self.dtype)
        abi.check(abi.lib().kg_grid_download(self._h, int(unbuffered), abi.ptr(out.reshape(-1))))
        return out.reshape(self.width, self.height)

    def step_stencil(self, rule=abi.KG_RULE_FOREST_FIRE):
        """one model step through the field API; `rule` = a shipped rule id, or the rule's body as a CUDA C
        expression in v, at(dx, dy), x, y, NONE (compiled at run time, include/krabgpu.h kg_grid_step_expr)"""
        if isinstance(rule, str):
            abi.check(abi.lib().kg_grid_step_expr(self._h, rule.encode()))
            return
        abi.check(abi.lib().kg_grid_step_stencil(self._h, rule))

    def run_stencil(self, nsteps, rule=abi.KG_RULE_FOREST_FIRE):
        abi.check(abi.lib().kg_grid_run_stencil(self._h, rule, nsteps))

    def init_forest_fire(self, density, seed):
        abi.check(abi.lib().kg_grid_init_forest_fire(self._h, density, seed))

    def sync(self):
        abi.check(abi.lib().kg_grid_sync(self._h))

    def run_stencil_timed(self, nsteps, rule=abi.KG_RULE_FOREST_FIRE):
        ms = C.c_double()
        abi.check(abi.lib().kg_grid_run_stencil_timed(self._h, rule, nsteps, C.byref(ms)))
        return ms.value

    def timer_start(self):
        abi.check(abi.lib().kg_grid_timer_start(self._h))

    def timer_stop(self):
        """milliseconds of device time since timer_start (CUDA events on the handle's stream)"""
        ms = C.c_double()
        abi.check(abi.lib().kg_grid_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def profile(self, enable=True):
        abi.check(abi.lib().kg_grid_profile(self._h, int(enable)))

    def profile_read(self, reset=True):
        ms = (C.c_double * 8)()
        ln = (abi.u64 * 8)()
        abi.check(abi.lib().kg_grid_profile_read(self._h, ms, ln, int(reset)))
        return {k: (ms[i], ln[i]) for i, k in enumerate(abi.KERNEL_KINDS)}
