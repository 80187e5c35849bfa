"""ctypes binding of include/krabgpu.h (libkrabgpu.so, hand-written CUDA for sm_100a).

This is the only way the Python host layer reaches the device.  There is no CPU fallback:
if the shared library is missing or no CUDA device is present, calls fail loudly.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# KRABGPU_LIB selects another build of the same sources (kernel experiments, tools/k4_ab.py)
_SO = os.environ.get("KRABGPU_LIB") or os.path.join(_PKG, "libkrabgpu.so")
_CSRC = os.path.join(_PKG, "csrc")

KG_OK, KG_E_CUDA, KG_E_INVALID, KG_E_CAPACITY, KG_E_OOB = 0, -1, -2, -3, -4
KG_BUF_READ, KG_BUF_WRITE = 0, 1
KG_QUERY_RELAX, KG_QUERY_EXACT = 0, 1
KG_ORDER_ANY, KG_ORDER_CANONICAL = 0, 1
KG_K4_AUTO, KG_K4_GENERIC, KG_K4_FAST_SCALAR, KG_K4_PACKED_BY_ID, KG_K4_TILED, KG_K4_COLTILE, KG_K4_STAGED = 0, 1, 2, 3, 4, 5, 6
KG_GRID_READ, KG_GRID_WRITE, KG_GRID_READWRITE = 0, 1, 2
KG_APPLY_CONST, KG_APPLY_ADD = 0, 1
KG_RULE_FOREST_FIRE = 0
KERNEL_KINDS = ("step", "hist", "scan", "scatter", "sortcell", "query", "misc", "stencil")


class KgError(RuntimeError):
    """A non-zero status from libkrabgpu (the Rust shim would panic! here)."""

    def __init__(self, code, msg):
        super().__init__(f"krabgpu error {code}: {msg}")
        self.code = code


class KgOutOfBounds(KgError):
    """KG_E_OOB: the reference would panic with an index out of bounds."""


class KgCustomStep(C.Structure):
    """include/krabgpu.h KgCustomStep"""
    _fields_ = [("pair", C.c_char_p), ("finish", C.c_char_p), ("consts", C.c_float * 16), ("nconsts", C.c_int32),
                ("radius", C.c_float), ("exact_query", C.c_int32), ("may_stop", C.c_int32), ("seed", C.c_uint64),
                ("step", C.c_uint64)]


class KgBoidsParams(C.Structure):
    _fields_ = [
        ("cohesion", C.c_float), ("avoidance", C.c_float), ("randomness", C.c_float),
        ("consistency", C.c_float), ("momentum", C.c_float), ("jump", C.c_float),
        ("radius", C.c_float), ("exact_query", C.c_int32), ("seed", C.c_uint64),
        ("step", C.c_uint64),
    ]


class KgLifeRule(C.Structure):
    _fields_ = [("death_prob", C.c_float), ("birth_prob", C.c_float), ("crowd_limit", C.c_uint32),
                ("reserved", C.c_uint32)]


def life_rule(death_prob=0.0, birth_prob=0.0, crowd_limit=0):
    """Dynamic population (include/krabgpu.h KgLifeRule): Agent::is_stopped / State::after_step births."""
    return KgLifeRule(death_prob, birth_prob, int(crowd_limit), 0)


def boids_params(radius=10.0, exact=0, seed=42, jump=0.7, cohesion=1.0, avoidance=1.0,
                 randomness=1.0, consistency=1.0, momentum=1.0, step=0):
    """Defaults are the fixture's constants (tests/model/flockers/bird.rs:12-17, :41)."""
    return KgBoidsParams(cohesion, avoidance, randomness, consistency, momentum, jump, radius,
                         int(exact), seed, step)


def build(force=False, verbose=False):
    """Compile libkrabgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)
            if f.endswith((".cu", ".cuh", "Makefile"))]
    srcs.append(os.path.join(os.path.dirname(_PKG), "include", "krabgpu.h"))
    fresh = os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs)
    if fresh and not force:
        return _SO
    cmd = ["make", "-C", _CSRC] + (["-B"] if force else [])
    out = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libkrabgpu.so failed")
    return _SO


_lib = None
vp = C.c_void_p
u64 = C.c_uint64
f32 = C.c_float
i32 = C.c_int32


def lib():
    """Load libkrabgpu.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise ImportError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C krabmaga_b200/csrc`; krabmaga_b200 has no CPU fallback")
    L = C.CDLL(_SO)
    P = C.POINTER
    sig = {
        "kg_last_error": (C.c_char_p, []),
        "kg_abi_version": (C.c_int, []),
        "kg_device_count": (C.c_int, []),
        "kg_launch_count": (u64, []),
        "kg_host_alloc": (C.c_int, [C.c_size_t, P(vp)]),
        "kg_host_free": (C.c_int, [vp]),
        "kg_field2d_create": (C.c_int, [f32, f32, f32, C.c_int, u64, C.c_int, P(vp)]),
        "kg_field2d_destroy": (C.c_int, [vp]),
        "kg_field2d_sync": (C.c_int, [vp]),
        "kg_field2d_dims": (C.c_int, [vp, P(i32), P(i32), P(i32), P(i32)]),
        "kg_field2d_set_order": (C.c_int, [vp, C.c_int]),
        "kg_field2d_set_kernel_variant": (C.c_int, [vp, C.c_int]),
        "kg_selftest_div": (C.c_int, [C.c_int, u64, u64, P(u64)]),
        "kg_field2d_set_object_locations": (C.c_int, [vp, u64, vp, vp, vp, vp, vp]),
        "kg_field2d_set_object_locations_dev": (C.c_int, [vp, u64, vp, vp, vp, vp, vp]),
        "kg_field2d_remove_object_location": (C.c_int, [vp, C.c_uint32, f32, f32]),
        "kg_field2d_lazy_update": (C.c_int, [vp]),
        "kg_field2d_update": (C.c_int, [vp]),
        "kg_field2d_nagents": (C.c_int, [vp, P(u64)]),
        "kg_field2d_num_objects": (C.c_int, [vp, C.c_int, P(u64)]),
        "kg_field2d_download": (C.c_int, [vp, C.c_int, u64, vp, vp, vp, vp, vp, vp, P(u64)]),
        "kg_field2d_cell_counts": (C.c_int, [vp, C.c_int, u64, vp]),
        "kg_field2d_num_objects_at_locations": (C.c_int, [vp, u64, vp, vp, vp]),
        "kg_field2d_get_objects": (C.c_int, [vp, C.c_int, f32, f32, u64, vp, P(u64)]),
        "kg_field2d_num_empty_bags": (C.c_int, [vp, P(u64)]),
        "kg_field2d_neighbors": (C.c_int, [vp, u64, vp, vp, f32, C.c_int, vp, vp, u64, P(u64)]),
        "kg_field2d_neighbors_agents": (C.c_int, [vp, u64, vp, vp, f32, C.c_int, vp, vp, vp, vp, vp, vp, u64,
                                                  P(u64)]),
        "kg_field2d_step_boids": (C.c_int, [vp, P(KgBoidsParams)]),
        "kg_field2d_run_boids": (C.c_int, [vp, P(KgBoidsParams), u64]),
        "kg_field2d_init_flockers": (C.c_int, [vp, u64, u64]),
        "kg_field2d_set_next_id": (C.c_int, [vp, C.c_uint32]),
        "kg_field2d_step_boids_life": (C.c_int, [vp, P(KgBoidsParams), P(KgLifeRule), P(u64), P(u64)]),
        "kg_field2d_step_boids_host": (C.c_int, [vp, P(KgBoidsParams), u64] + [vp] * 10),
        "kg_field2d_step_boids_host_ordered": (C.c_int, [vp, P(KgBoidsParams), u64] + [vp] * 9),
        "kg_field2d_reduce": (C.c_int, [vp, vp]),
        "kg_field2d_step_custom": (C.c_int, [vp, vp]),
        "kg_jit_agent_source": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, vp, u64, P(u64)]),
        "kg_field2d_run_boids_series": (C.c_int, [vp, P(KgBoidsParams), u64, u64, vp, u64]),
        "kg_field2d_l2_flush": (C.c_int, [vp, u64]),
        "kg_field2d_run_boids_timed": (C.c_int, [vp, P(KgBoidsParams), u64, u64, P(C.c_double)]),
        "kg_field2d_timer_start": (C.c_int, [vp]),
        "kg_field2d_timer_stop": (C.c_int, [vp, P(C.c_double)]),
        "kg_field2d_profile": (C.c_int, [vp, C.c_int]),
        "kg_field2d_profile_read": (C.c_int, [vp, P(C.c_double), P(u64), C.c_int]),
        "kg_strip_create": (C.c_int, [f32, f32, f32, C.c_int, f32, C.c_int, C.c_int, u64, u64, u64,
                                      C.c_int, P(vp)]),
        "kg_strip_destroy": (C.c_int, [vp]),
        "kg_strip_columns": (C.c_int, [vp, P(i32), P(i32), P(i32), P(i32), P(i32)]),
        "kg_strip_set_order": (C.c_int, [vp, C.c_int]),
        "kg_strip_ipc_export": (C.c_int, [vp, vp]),
        "kg_strip_connect_ipc": (C.c_int, [vp, vp, vp]),
        "kg_strip_connect_local": (C.c_int, [vp, vp, vp]),
        "kg_strip_init_flockers": (C.c_int, [vp, u64, u64]),
        "kg_strip_upload": (C.c_int, [vp, u64, vp, vp, vp, vp, vp]),
        "kg_strip_clear": (C.c_int, [vp]),
        "kg_strip_prepare": (C.c_int, [vp]),
        "kg_strip_step_boids": (C.c_int, [vp, P(KgBoidsParams)]),
        "kg_strip_run_boids": (C.c_int, [vp, P(KgBoidsParams), u64]),
        "kg_strip_run_boids_timed": (C.c_int, [vp, P(KgBoidsParams), u64, u64, P(C.c_double)]),
        "kg_strip_sync": (C.c_int, [vp]),
        "kg_strip_stats": (C.c_int, [vp] + [P(u64)] * 6),
        "kg_strip_download": (C.c_int, [vp, u64, vp, vp, vp, vp, vp, P(u64)]),
        "kg_strip_timer_start": (C.c_int, [vp]),
        "kg_strip_timer_stop": (C.c_int, [vp, P(C.c_double)]),
        "kg_gridstrip_create": (C.c_int, [i32, i32, C.c_int, C.c_int, C.c_int, P(vp)]),
        "kg_gridstrip_destroy": (C.c_int, [vp]),
        "kg_gridstrip_rows": (C.c_int, [vp, P(i32), P(i32)]),
        "kg_gridstrip_pass_plan": (C.c_int, [C.c_int32, C.c_int32, C.c_int, u64, vp, vp, u64, P(u64)]),
        "kg_gridstrip_ipc_export": (C.c_int, [vp, vp]),
        "kg_gridstrip_connect_ipc": (C.c_int, [vp, vp, vp]),
        "kg_gridstrip_connect_local": (C.c_int, [vp, vp, vp]),
        "kg_gridstrip_init_forest_fire": (C.c_int, [vp, f32, u64]),
        "kg_gridstrip_upload": (C.c_int, [vp, vp]),
        "kg_gridstrip_download": (C.c_int, [vp, vp]),
        "kg_gridstrip_prepare": (C.c_int, [vp]),
        "kg_gridstrip_run_stencil": (C.c_int, [vp, C.c_int, u64]),
        "kg_gridstrip_run_stencil_timed": (C.c_int, [vp, C.c_int, u64, P(C.c_double)]),
        "kg_gridstrip_sync": (C.c_int, [vp]),
        "kg_batch_create": (C.c_int, [f32, f32, f32, C.c_int, C.c_uint32, C.c_uint32, C.c_int, P(vp)]),
        "kg_batch_destroy": (C.c_int, [vp]),
        "kg_batch_dims": (C.c_int, [vp, P(C.c_uint32), P(C.c_uint32), P(i32), P(i32)]),
        "kg_batch_set_order": (C.c_int, [vp, C.c_int]),
        "kg_batch_set_params": (C.c_int, [vp, C.c_uint32, C.c_uint32, P(KgBoidsParams)]),
        "kg_batch_init_flockers": (C.c_int, [vp]),
        "kg_batch_upload": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "kg_batch_lazy_update": (C.c_int, [vp]),
        "kg_batch_step_boids": (C.c_int, [vp, u64]),
        "kg_batch_run_boids": (C.c_int, [vp, u64, u64]),
        "kg_batch_run_boids_timed": (C.c_int, [vp, u64, u64, u64, P(C.c_double)]),
        "kg_batch_download": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
        "kg_batch_reduce": (C.c_int, [vp, vp]),
        "kg_batch_sync": (C.c_int, [vp]),
        "kg_batch_timer_start": (C.c_int, [vp]),
        "kg_batch_timer_stop": (C.c_int, [vp, P(C.c_double)]),
        "kg_block_create": (C.c_int, [C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                      C.c_int, u64, u64, C.c_int, P(vp)]),
        "kg_block_destroy": (C.c_int, [vp]),
        "kg_block_cells": (C.c_int, [vp, vp, vp]),
        "kg_block_set_order": (C.c_int, [vp, C.c_int]),
        "kg_block_upload": (C.c_int, [vp, u64, vp, vp, vp, vp, vp]),
        "kg_block_lazy_update": (C.c_int, [vp]),
        "kg_blocks_step": (C.c_int, [P(vp), C.c_int, P(KgBoidsParams)]),
        "kg_blocks_run": (C.c_int, [P(vp), C.c_int, P(KgBoidsParams), u64]),
        "kg_block_download": (C.c_int, [vp, u64, vp, vp, vp, vp, vp, P(u64)]),
        "kg_block_counts": (C.c_int, [vp, P(u64), P(u64)]),
        "kg_block_partition": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, P(i32)]),
        "kg_objgrid_create": (C.c_int, [i32, i32, u64, C.c_int, P(vp)]),
        "kg_objgrid_create_sparse": (C.c_int, [i32, i32, u64, C.c_int, P(vp)]),
        "kg_objgrid_destroy": (C.c_int, [vp]),
        "kg_objgrid_dims": (C.c_int, [vp, P(i32), P(i32), P(u64)]),
        "kg_objgrid_set_object_locations": (C.c_int, [vp, u64, vp, vp, vp, vp]),
        "kg_objgrid_remove_object_locations": (C.c_int, [vp, u64, vp, vp, vp]),
        "kg_objgrid_lazy_update": (C.c_int, [vp]),
        "kg_objgrid_update": (C.c_int, [vp]),
        "kg_objgrid_num_objects": (C.c_int, [vp, C.c_int, P(u64)]),
        "kg_objgrid_get_objects": (C.c_int, [vp, C.c_int, i32, i32, u64, vp, vp, P(u64)]),
        "kg_objgrid_get_location": (C.c_int, [vp, C.c_int, C.c_uint32, P(i32), P(i32), P(C.c_int)]),
        "kg_objgrid_iter_objects": (C.c_int, [vp, C.c_int, u64, vp, vp, vp, vp, P(u64)]),
        "kg_objgrid_bag_sizes": (C.c_int, [vp, C.c_int, u64, vp]),
        "kg_objgrid_apply": (C.c_int, [vp, C.c_int, C.c_uint32, C.c_int, P(u64)]),
        "kg_grid_create": (C.c_int, [i32, i32, C.c_int, C.c_uint32, C.c_int, P(vp)]),
        "kg_grid_destroy": (C.c_int, [vp]),
        "kg_grid_sync": (C.c_int, [vp]),
        "kg_grid_set_values": (C.c_int, [vp, u64, vp, vp, vp]),
        "kg_grid_remove_values": (C.c_int, [vp, u64, vp, vp]),
        "kg_grid_get_values": (C.c_int, [vp, C.c_int, u64, vp, vp, vp]),
        "kg_grid_upload": (C.c_int, [vp, C.c_int, vp]),
        "kg_grid_download": (C.c_int, [vp, C.c_int, vp]),
        "kg_grid_apply": (C.c_int, [vp, C.c_int, C.c_uint32, C.c_int]),
        "kg_grid_apply_expr": (C.c_int, [vp, C.c_char_p, C.c_int]),
        "kg_grid_step_expr": (C.c_int, [vp, C.c_char_p]),
        "kg_grid_get_location": (C.c_int, [vp, C.c_int, C.c_uint32, P(i32), P(i32), P(C.c_int)]),
        "kg_grid_num_empty": (C.c_int, [vp, P(u64)]),
        "kg_grid_lazy_update": (C.c_int, [vp]),
        "kg_grid_update": (C.c_int, [vp]),
        "kg_grid_step_stencil": (C.c_int, [vp, C.c_int]),
        "kg_grid_run_stencil": (C.c_int, [vp, C.c_int, u64]),
        "kg_grid_init_forest_fire": (C.c_int, [vp, f32, u64]),
        "kg_grid_run_stencil_timed": (C.c_int, [vp, C.c_int, u64, P(C.c_double)]),
        "kg_grid_timer_start": (C.c_int, [vp]),
        "kg_grid_timer_stop": (C.c_int, [vp, P(C.c_double)]),
        "kg_grid_profile": (C.c_int, [vp, C.c_int]),
        "kg_grid_profile_read": (C.c_int, [vp, P(C.c_double), P(u64), C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._declared = sorted(sig)
    _lib = L
    return L


def check(rc):
    if rc == KG_OK:
        return
    msg = lib().kg_last_error().decode(errors="replace")
    if rc == KG_E_OOB:
        raise KgOutOfBounds(rc, msg)
    raise KgError(rc, msg)


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(vp)


def as_f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def as_u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def pinned_empty(n, dtype):
    """numpy array over page-locked host memory from kg_host_alloc; the memory is returned with
    kg_host_free when the array (and every view of it) has been garbage collected."""
    import weakref
    dtype = np.dtype(dtype)
    p = vp()
    check(lib().kg_host_alloc(max(1, n * dtype.itemsize), C.byref(p)))
    buf = (C.c_char * (n * dtype.itemsize)).from_address(p.value)
    # views made by numpy keep `buf` alive through .base, so its finaliser runs after the last of them
    weakref.finalize(buf, lib().kg_host_free, p.value)
    return np.frombuffer(buf, dtype=dtype, count=n)
