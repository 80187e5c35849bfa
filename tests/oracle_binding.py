"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference (oracle/krabmaga_oracle.hpp).  It is the
checker for the parity tests and the CPU baseline of bench.py; the product package
`krabmaga_b200` never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "liboracle.so")

u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


class OkgBoidsParams(C.Structure):
    _fields_ = [
        ("cohesion", C.c_float), ("avoidance", C.c_float), ("randomness", C.c_float),
        ("consistency", C.c_float), ("momentum", C.c_float), ("jump", C.c_float),
        ("radius", C.c_float), ("exact_query", C.c_int32), ("seed", C.c_uint64),
        ("step", C.c_uint64),
    ]


def boids_params(radius=10.0, exact=0, seed=42, jump=0.7, cohesion=1.0, avoidance=1.0,
                 randomness=1.0, consistency=1.0, momentum=1.0, step=0):
    return OkgBoidsParams(cohesion, avoidance, randomness, consistency, momentum, jump, radius,
                          int(exact), seed, step)


def build():
    src = [os.path.join(_ROOT, "oracle", f) for f in
           ("oracle_capi.cpp", "krabmaga_oracle.hpp", "object_grid.hpp", "philox.hpp", "Makefile")]
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src):
        return _SO
    subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp = C.c_void_p
    sig = {
        "okg_last_error": (C.c_char_p, []),
        "okg_philox4x32_10": (None, [u32p, u32p, u32p]),
        "okg_u01_f32": (C.c_float, [C.c_uint32]),
        "okg_toroidal_transform": (C.c_float, [C.c_float, C.c_float]),
        "okg_toroidal_distance": (C.c_float, [C.c_float, C.c_float, C.c_float]),
        "okg_t_transform": (C.c_int, [C.c_int, C.c_int]),
        "okg_field2d_new": (vp, [C.c_float, C.c_float, C.c_float, C.c_int]),
        "okg_field2d_free": (None, [vp]),
        "okg_field2d_dims": (None, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "okg_field2d_nagents": (C.c_uint64, [vp]),
        "okg_field2d_discretize": (C.c_int, [vp, C.c_float, C.c_float, C.POINTER(C.c_int),
                                             C.POINTER(C.c_int)]),
        "okg_field2d_set_object_location": (C.c_int, [vp, C.c_uint32] + [C.c_float] * 4),
        "okg_field2d_set_object_locations": (C.c_int, [vp, C.c_uint64, u32p, f32p, f32p, f32p, f32p]),
        "okg_field2d_remove_object_location": (C.c_int, [vp, C.c_uint32, C.c_float, C.c_float]),
        "okg_field2d_lazy_update": (None, [vp]),
        "okg_field2d_update": (None, [vp]),
        "okg_field2d_neighbors": (C.c_int64, [vp, C.c_float, C.c_float, C.c_float, C.c_int, u32p,
                                              C.c_uint64]),
        "okg_field2d_neighbors_batch": (C.c_int, [vp, C.c_uint64, f32p, f32p, C.c_float, C.c_int,
                                                  u64p, u32p, C.c_uint64]),
        "okg_field2d_get_objects": (C.c_int64, [vp, C.c_float, C.c_float, C.c_int, u32p, C.c_uint64]),
        "okg_field2d_num_objects_at_location": (C.c_int64, [vp, C.c_float, C.c_float]),
        "okg_field2d_get_empty_bags": (C.c_int64, [vp, f32p, f32p, C.c_uint64]),
        "okg_field2d_iter_objects": (C.c_int64, [vp, C.c_int, C.c_uint64] + [vp] * 8),
        "okg_field2d_cell_counts": (C.c_int64, [vp, C.c_int, u32p, C.c_uint64]),
        "okg_flockers_new": (vp, [C.c_float, C.c_float, C.c_float, C.c_int, C.c_uint32,
                                  C.POINTER(OkgBoidsParams), C.c_int]),
        "okg_flockers_free": (None, [vp]),
        "okg_flockers_preset": (None, [vp, C.c_uint64, u32p, f32p, f32p, f32p, f32p]),
        "okg_flockers_init": (C.c_int, [vp]),
        "okg_flockers_step": (C.c_int, [vp, C.c_uint64]),
        "okg_flockers_schedule_step": (C.c_uint64, [vp]),
        "okg_flockers_field": (vp, [vp]),
        "okg_flockers_agents": (C.c_int, [vp, C.c_uint64, f32p, f32p, f32p, f32p]),
        "okg_flockers_set_life": (None, [vp, C.c_float, C.c_float, C.c_uint32, C.c_uint32]),
        "okg_flockers_population": (C.c_uint64, [vp, C.c_uint64, u32p, f32p, f32p, f32p, f32p,
                                                 C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "okg_flockers_pop_order": (C.c_int, [vp, u32p, C.c_uint64]),
        "okg_flockers_time_steps": (C.c_double, [vp, C.c_uint64]),
        "okg_flockers_sweep": (C.c_double, [C.c_float, C.c_float, C.c_float, C.c_int, C.c_uint32,
                                            C.POINTER(OkgBoidsParams), C.c_uint32, C.c_uint64,
                                            C.c_uint32, C.POINTER(C.c_uint64)]),
        "okg_schedule_new": (vp, []),
        "okg_schedule_free": (None, [vp]),
        "okg_schedule_repeating": (C.c_int, [vp, C.c_uint32, C.c_float, C.c_int,
                                             C.POINTER(C.c_uint32)]),
        "okg_schedule_events": (C.c_int64, [vp, u32p, C.c_uint64]),
        "okg_schedule_dequeue": (C.c_int, [vp, C.c_uint32]),
        "okg_grid_new": (vp, [C.c_int, C.c_int]),
        "okg_grid_free": (None, [vp]),
        "okg_grid_set_value_location": (C.c_int, [vp, C.c_uint16, C.c_int, C.c_int]),
        "okg_grid_remove_value_location": (C.c_int, [vp, C.c_int, C.c_int]),
        "okg_grid_get_value": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint16)]),
        "okg_grid_get_location": (C.c_int, [vp, C.c_uint16, C.c_int, C.POINTER(C.c_int),
                                            C.POINTER(C.c_int)]),
        "okg_grid_num_empty_bags": (C.c_int64, [vp]),
        "okg_grid_lazy_update": (None, [vp]),
        "okg_grid_update": (None, [vp]),
        "okg_grid_apply": (None, [vp, C.c_int, C.c_uint16, C.c_int]),
        "okg_grid_dump": (None, [vp, C.c_int, C.c_uint16,
                                 np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")]),
        "okg_ff_new": (vp, [C.c_int, C.c_int]),
        "okg_ff_free": (None, [vp]),
        "okg_ff_init": (None, [vp, C.c_float, C.c_uint64]),
        "okg_ff_load": (None, [vp, np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS"), C.c_uint8]),
        "okg_ff_step": (None, [vp, C.c_uint64]),
        "okg_ff_time_steps": (C.c_double, [vp, C.c_uint64]),
        "okg_ff_dump": (None, [vp, C.c_uint8, np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")]),
        "okg_hardware_concurrency": (C.c_uint, []),
        "okg_ogrid_new": (vp, [C.c_int, C.c_int]),
        "okg_ogrid_free": (None, [vp]),
        "okg_ogrid_set_object_location": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int]),
        "okg_ogrid_remove_object_location": (C.c_int, [vp, C.c_uint32, C.c_int, C.c_int]),
        "okg_ogrid_lazy_update": (C.c_int, [vp]),
        "okg_ogrid_update": (C.c_int, [vp]),
        "okg_ogrid_nbags": (C.c_uint64, [vp, C.c_int]),
        "okg_ogrid_get_objects": (C.c_int64, [vp, C.c_int, C.c_int, C.c_int, u32p, u32p, C.c_uint64]),
        "okg_ogrid_get_location": (C.c_int, [vp, C.c_int, C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "okg_ogrid_get_empty_bags": (C.c_int64, [vp, i32p, i32p, C.c_uint64]),
        "okg_ogrid_iter_objects": (C.c_int64, [vp, C.c_int, i32p, i32p, u32p, u32p, C.c_uint64]),
        "okg_ogrid_apply": (C.c_int64, [vp, C.c_int, C.c_uint32, C.c_int]),
        "okg_sgrid_new": (vp, [C.c_int, C.c_int]),
        "okg_sgrid_free": (None, [vp]),
        "okg_sgrid_set_object_location": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int]),
        "okg_sgrid_remove_object_location": (C.c_int, [vp, C.c_uint32, C.c_int, C.c_int]),
        "okg_sgrid_lazy_update": (C.c_int, [vp]),
        "okg_sgrid_update": (C.c_int, [vp]),
        "okg_sgrid_get_objects": (C.c_int64, [vp, C.c_int, C.c_int, C.c_int, u32p, u32p, C.c_uint64]),
        "okg_sgrid_get_location": (C.c_int, [vp, C.c_int, C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "okg_sgrid_get_empty_bags": (C.c_int64, [vp, i32p, i32p, C.c_uint64]),
        "okg_sgrid_iter_objects": (C.c_int64, [vp, C.c_int, i32p, i32p, u32p, u32p, C.c_uint64]),
        "okg_sgrid_apply": (C.c_int64, [vp, C.c_int, C.c_uint32, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class OraclePanic(RuntimeError):
    """A restated Rust panic (index out of bounds etc.)."""


def _check(rc):
    if rc < 0:
        raise OraclePanic(lib().okg_last_error().decode())
    return rc


def philox(ctr, key):
    out = np.zeros(4, np.uint32)
    lib().okg_philox4x32_10(np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), out)
    return out


class Field2D:
    """Restated `Field2D<Bird>` (field_2d.rs:269-921)."""

    def __init__(self, w, h, d, toroidal, _borrowed=None):
        self._own = _borrowed is None
        self.p = lib().okg_field2d_new(w, h, d, int(toroidal)) if self._own else _borrowed
        self.width, self.height, self.discretization, self.toroidal = w, h, d, toroidal

    def __del__(self):
        if getattr(self, "_own", False) and self.p:
            lib().okg_field2d_free(self.p)
            self.p = None

    def dims(self):
        dw, dh, nr, nw = C.c_int(), C.c_int(), C.c_uint64(), C.c_uint64()
        lib().okg_field2d_dims(self.p, dw, dh, nr, nw)
        return dw.value, dh.value, nr.value, nw.value

    @property
    def nagents(self):
        return lib().okg_field2d_nagents(self.p)

    def discretize(self, x, y):
        cx, cy = C.c_int(), C.c_int()
        lib().okg_field2d_discretize(self.p, x, y, cx, cy)
        return cx.value, cy.value

    def set_object_location(self, id, x, y, ldx=0.0, ldy=0.0):
        _check(lib().okg_field2d_set_object_location(self.p, id, x, y, ldx, ldy))

    def set_object_locations(self, ids, x, y, ldx, ldy):
        a = [np.ascontiguousarray(ids, np.uint32)] + [np.ascontiguousarray(v, np.float32)
                                                       for v in (x, y, ldx, ldy)]
        _check(lib().okg_field2d_set_object_locations(self.p, len(a[0]), *a))

    def remove_object_location(self, id, x, y):
        _check(lib().okg_field2d_remove_object_location(self.p, id, x, y))

    def lazy_update(self):
        lib().okg_field2d_lazy_update(self.p)

    def update(self):
        lib().okg_field2d_update(self.p)

    def _neigh(self, x, y, dist, mode):
        cap = 1024
        while True:
            out = np.zeros(cap, np.uint32)
            n = _check(lib().okg_field2d_neighbors(self.p, x, y, dist, mode, out, cap))
            if n <= cap:
                return out[:n].copy()
            cap = n

    def get_neighbors_within_distance(self, x, y, dist):
        return self._neigh(x, y, dist, 1)

    def get_neighbors_within_relax_distance(self, x, y, dist):
        return self._neigh(x, y, dist, 0)

    def neighbors_batch(self, qx, qy, dist, mode):
        qx = np.ascontiguousarray(qx, np.float32)
        qy = np.ascontiguousarray(qy, np.float32)
        offs = np.zeros(len(qx) + 1, np.uint64)
        cap = max(1024, 64 * len(qx))
        while True:
            ids = np.zeros(cap, np.uint32)
            _check(lib().okg_field2d_neighbors_batch(self.p, len(qx), qx, qy, dist, mode, offs, ids,
                                                     cap))
            if offs[-1] <= cap:
                return offs.astype(np.int64), ids[: int(offs[-1])].copy()
            cap = int(offs[-1])

    def get_objects(self, x, y, unbuffered=False):
        out = np.zeros(4096, np.uint32)
        n = _check(lib().okg_field2d_get_objects(self.p, x, y, int(unbuffered), out, len(out)))
        return out[:n].copy()

    def get_objects_unbuffered(self, x, y):
        return self.get_objects(x, y, True)

    def num_objects_at_location(self, x, y):
        return _check(lib().okg_field2d_num_objects_at_location(self.p, x, y))

    def get_empty_bags(self):
        dw, dh, _, _ = self.dims()
        ox = np.zeros(dw * dh, np.float32)
        oy = np.zeros(dw * dh, np.float32)
        n = _check(lib().okg_field2d_get_empty_bags(self.p, ox, oy, len(ox)))
        return np.stack([ox[:n], oy[:n]], 1)

    def iter_objects(self, unbuffered=False):
        """dict of arrays in iteration order: id,x,y,ldx,ldy,cell,ox,oy"""
        n = _check(lib().okg_field2d_iter_objects(self.p, int(unbuffered), 0, *([None] * 8)))
        arrs = dict(id=np.zeros(n, np.uint32), x=np.zeros(n, np.float32), y=np.zeros(n, np.float32),
                    ldx=np.zeros(n, np.float32), ldy=np.zeros(n, np.float32),
                    cell=np.zeros(n, np.int32), ox=np.zeros(n, np.float32),
                    oy=np.zeros(n, np.float32))
        ptrs = [a.ctypes.data_as(C.c_void_p) for a in arrs.values()]
        _check(lib().okg_field2d_iter_objects(self.p, int(unbuffered), n, *ptrs))
        return arrs

    def cell_counts(self, unbuffered=False):
        dw, dh, nr, nw = self.dims()
        out = np.zeros(max(nr, nw), np.uint32)
        n = lib().okg_field2d_cell_counts(self.p, int(unbuffered), out, len(out))
        return out[:n].copy()


class Flockers:
    """Restated Flockers fixture (tests/model/flockers) + sequential Schedule."""

    def __init__(self, w, h, n, disc, toroidal=True, params=None, canonical_order=False):
        self.params = params or boids_params()
        self.n = n
        self.p = lib().okg_flockers_new(w, h, disc, int(toroidal), n, C.byref(self.params),
                                        int(canonical_order))
        self.field1 = Field2D(w, h, disc, toroidal, _borrowed=lib().okg_flockers_field(self.p))

    def __del__(self):
        if getattr(self, "p", None):
            lib().okg_flockers_free(self.p)
            self.p = None

    def preset(self, ids, x, y, ldx, ldy):
        a = [np.ascontiguousarray(ids, np.uint32)] + [np.ascontiguousarray(v, np.float32)
                                                       for v in (x, y, ldx, ldy)]
        self.n = len(a[0])
        lib().okg_flockers_preset(self.p, self.n, *a)

    def init(self):
        _check(lib().okg_flockers_init(self.p))

    def step(self, nsteps=1):
        _check(lib().okg_flockers_step(self.p, nsteps))

    @property
    def schedule_step(self):
        return lib().okg_flockers_schedule_step(self.p)

    def set_life(self, death_prob, birth_prob, crowd_limit, next_id):
        """dynamic population (oracle LifeRule): call before init()"""
        lib().okg_flockers_set_life(self.p, death_prob, birth_prob, crowd_limit, next_id)

    def population(self):
        """dict(id, x, y, ldx, ldy) of the scheduled agents sorted by id, plus born / died totals"""
        born, died = C.c_uint64(), C.c_uint64()
        z, zf = np.zeros(1, np.uint32), np.zeros(1, np.float32)
        n = lib().okg_flockers_population(self.p, 0, z, zf, zf, zf, zf, None, None)
        ids = np.zeros(max(n, 1), np.uint32)
        a = [np.zeros(max(n, 1), np.float32) for _ in range(4)]
        lib().okg_flockers_population(self.p, n, ids, *a, C.byref(born), C.byref(died))
        ids, a = ids[:n], [v[:n] for v in a]
        o = np.argsort(ids, kind="stable")
        return dict(id=ids[o], x=a[0][o], y=a[1][o], ldx=a[2][o], ldy=a[3][o], born=born.value, died=died.value)

    def agents(self):
        """(x, y, ldx, ldy) of every agent, indexed by id (ids must be 0..n-1)."""
        out = [np.zeros(self.n, np.float32) for _ in range(4)]
        _check(lib().okg_flockers_agents(self.p, self.n, *out))
        return out

    def pop_order(self):
        out = np.zeros(self.n, np.uint32)
        _check(lib().okg_flockers_pop_order(self.p, out, self.n))
        return out

    def time_steps(self, nsteps):
        return lib().okg_flockers_time_steps(self.p, nsteps)


def flockers_sweep(w, h, disc, toroidal, n, params, replicas, nsteps, threads=0):
    work = C.c_uint64()
    sec = lib().okg_flockers_sweep(w, h, disc, int(toroidal), n, C.byref(params), replicas, nsteps,
                                   threads, C.byref(work))
    return sec, work.value


class Schedule:
    def __init__(self):
        self.p = lib().okg_schedule_new()

    def __del__(self):
        if getattr(self, "p", None):
            lib().okg_schedule_free(self.p)
            self.p = None

    def schedule_repeating(self, tag, time=0.0, ordering=0):
        ido = C.c_uint32()
        ok = lib().okg_schedule_repeating(self.p, tag, time, ordering, C.byref(ido))
        return ido.value, bool(ok)

    def get_all_events(self):
        out = np.zeros(1 << 16, np.uint32)
        n = lib().okg_schedule_events(self.p, out, len(out))
        return out[:n].copy()

    def dequeue(self, id):
        return bool(lib().okg_schedule_dequeue(self.p, id))


class DenseNumberGrid2D:
    """Restated `DenseNumberGrid2D<u16>` (dense_number_grid_2d.rs:90-561)."""
    READ, WRITE, READWRITE = 0, 1, 2

    def __init__(self, w, h):
        self.p = lib().okg_grid_new(w, h)
        if not self.p:
            raise OraclePanic(lib().okg_last_error().decode())
        self.width, self.height = abs(w), abs(h)

    def __del__(self):
        if getattr(self, "p", None):
            lib().okg_grid_free(self.p)
            self.p = None

    def set_value_location(self, v, x, y):
        _check(lib().okg_grid_set_value_location(self.p, v, x, y))

    def remove_value_location(self, x, y):
        _check(lib().okg_grid_remove_value_location(self.p, x, y))

    def get_value(self, x, y, unbuffered=False):
        v = C.c_uint16()
        rc = _check(lib().okg_grid_get_value(self.p, x, y, int(unbuffered), C.byref(v)))
        return v.value if rc == 1 else None

    def get_value_unbuffered(self, x, y):
        return self.get_value(x, y, True)

    def get_location(self, v, unbuffered=False):
        x, y = C.c_int(), C.c_int()
        rc = lib().okg_grid_get_location(self.p, v, int(unbuffered), C.byref(x), C.byref(y))
        return (x.value, y.value) if rc == 1 else None

    def get_location_unbuffered(self, v):
        return self.get_location(v, True)

    def num_empty_bags(self):
        return lib().okg_grid_num_empty_bags(self.p)

    def lazy_update(self):
        lib().okg_grid_lazy_update(self.p)

    def update(self):
        lib().okg_grid_update(self.p)

    def apply_const(self, c, option):
        lib().okg_grid_apply(self.p, 0, c, option)

    def apply_add(self, c, option):
        lib().okg_grid_apply(self.p, 1, c, option)

    def dump(self, unbuffered=False, none=0xFFFF):
        out = np.zeros(self.width * self.height, np.uint16)
        lib().okg_grid_dump(self.p, int(unbuffered), none, out)
        return out.reshape(self.width, self.height)


class ForestFire:
    GREEN, BURNING, BURNED, NONE = 1, 2, 3, 0xFF

    def __init__(self, w, h):
        self.p = lib().okg_ff_new(w, h)
        self.w, self.h = w, h

    def __del__(self):
        if getattr(self, "p", None):
            lib().okg_ff_free(self.p)
            self.p = None

    def init(self, density, seed):
        lib().okg_ff_init(self.p, density, seed)

    def load(self, cells, none=0xFF):
        lib().okg_ff_load(self.p, np.ascontiguousarray(cells, np.uint8).reshape(-1), none)

    def step(self, n=1):
        lib().okg_ff_step(self.p, n)

    def time_steps(self, n):
        return lib().okg_ff_time_steps(self.p, n)

    def dump(self, none=0xFF):
        out = np.zeros(self.w * self.h, np.uint8)
        lib().okg_ff_dump(self.p, none, out)
        return out.reshape(self.w, self.h)


class DenseGrid2D:
    """oracle::DenseGrid2D<GridObj> (dense_object_grid_2d.rs:175-779).  Objects are (id, tag) pairs
    that compare by id, like the fixture's Bird; the tag stands in for Bird.flag."""
    READ, WRITE, READWRITE = 0, 1, 2
    SET_TAG, REMOVE, REMOVE_IF_TAG, TAG_WITH_BAG_ID = 0, 1, 2, 3

    _pfx = "okg_ogrid_"

    def _call(self, name, *args):
        return getattr(lib(), self._pfx + name)(*args)

    def __init__(self, width, height):
        self.p = self._call("new", width, height)
        if not self.p:
            raise OraclePanic(lib().okg_last_error().decode())
        self.width, self.height = abs(width), abs(height)

    def __del__(self):
        if getattr(self, "p", None):
            self._call("free", self.p)
            self.p = None

    def set_object_location(self, obj, loc):
        _check(self._call("set_object_location", self.p, obj[0], obj[1], loc[0], loc[1]))

    def remove_object_location(self, obj, loc):
        _check(self._call("remove_object_location", self.p, obj[0], loc[0], loc[1]))

    def lazy_update(self):
        _check(self._call("lazy_update", self.p))

    def update(self):
        _check(self._call("update", self.p))

    def nbags(self, unbuffered=False):
        return int(self._call("nbags", self.p, int(unbuffered)))

    def get_objects(self, loc, unbuffered=False):
        """list of (id, tag), or None for an empty bag"""
        cap = 64
        while True:
            ids, tags = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
            n = self._call("get_objects", self.p, int(unbuffered), loc[0], loc[1], ids, tags, cap)
            if n == -2:
                return None
            _check(n)
            if n <= cap:
                return [(int(a), int(b)) for a, b in zip(ids[:n], tags[:n])]
            cap = int(n)

    def get_objects_unbuffered(self, loc):
        return self.get_objects(loc, True)

    def get_location(self, obj, unbuffered=False):
        x, y = C.c_int(), C.c_int()
        r = _check(self._call("get_location", self.p, int(unbuffered), obj[0], C.byref(x), C.byref(y)))
        return (x.value, y.value) if r else None

    def get_location_unbuffered(self, obj):
        return self.get_location(obj, True)

    def get_empty_bags(self):
        cap = max(self.nbags(), 1)
        xs, ys = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        n = _check(self._call("get_empty_bags", self.p, xs, ys, cap))
        return [(int(a), int(b)) for a, b in zip(xs[:n], ys[:n])]

    def iter_objects(self, unbuffered=False):
        """[((x, y), (id, tag)), ...] in closure-call order"""
        cap = 1024
        while True:
            xs, ys = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
            ids, tags = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
            n = _check(self._call("iter_objects", self.p, int(unbuffered), xs, ys, ids, tags, cap))
            if n <= cap:
                return [((int(xs[i]), int(ys[i])), (int(ids[i]), int(tags[i]))) for i in range(n)]
            cap = int(n)

    def iter_objects_unbuffered(self):
        return self.iter_objects(True)

    def apply_to_all_values(self, op, arg, option):
        """closure family of okg_ogrid_apply; returns the number of closure calls"""
        return _check(self._call("apply", self.p, op, arg, option))


class SparseGrid2D(DenseGrid2D):
    """oracle::SparseGrid2D<GridObj> (sparse_object_grid_2d.rs:203-721): same verbs over two hash
    maps; iteration order is unspecified, compare as sets."""
    _pfx = "okg_sgrid_"

    def nbags(self, unbuffered=False):
        return self.width * self.height
