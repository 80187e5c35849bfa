"""GPU `Field2D` — the sibling of the reference's two Field2D variants
(src/engine/fields/field_2d.rs:28-266 DBDashMap, :267-921 default) that a model opts into by
swapping the field type.  Same method names and argument meaning; agents are the Flockers `Bird`
payload (id, pos, last_d).  Failures raise (the reference panics); out-of-grid coordinates raise
KgOutOfBounds where the reference's Vec index panics.

Every method is one call across the C ABI (include/krabgpu.h) — batch variants take arrays so the
boundary is crossed once per step, not once per agent.
"""
import ctypes as C

import random

import numpy as np

from ... import _abi as abi
from ..location import Real2D
from .field import Field


class Field2D(Field):
    def __init__(self, w, h, d, t, capacity=1 << 20, device=0):
        """Field2D::new(w, h, d, t)  field_2d.rs:304-322 (+ device capacity)."""
        self._h = abi.vp()
        abi.check(abi.lib().kg_field2d_create(w, h, d, int(bool(t)), capacity, device, C.byref(self._h)))
        self.width, self.height, self.discretization, self.toroidal = w, h, d, bool(t)
        self.capacity, self.device = capacity, device
        dw, dh, mx, my = abi.i32(), abi.i32(), abi.i32(), abi.i32()
        abi.check(abi.lib().kg_field2d_dims(self._h, dw, dh, mx, my))
        self.dw, self.dh, self.max_x, self.max_y = dw.value, dh.value, mx.value, my.value

    def close(self):
        if getattr(self, "_h", None):
            abi.lib().kg_field2d_destroy(self._h)
            self._h = None

    __del__ = close

    # ------------------------------------------------------------------ write side
    def set_object_location(self, object, loc):
        """field_2d.rs:838-846.  `object` = (id, last_dx, last_dy) or an object with .id/.last_d."""
        oid, ldx, ldy = _unpack(object)
        self.set_object_locations([oid], [loc[0]], [loc[1]], [ldx], [ldy])

    def set_object_locations(self, ids, x, y, last_dx=None, last_dy=None):
        ids, x, y = abi.as_u32(ids), abi.as_f32(x), abi.as_f32(y)
        n = len(ids)
        ldx = abi.as_f32(np.zeros(n) if last_dx is None else last_dx)
        ldy = abi.as_f32(np.zeros(n) if last_dy is None else last_dy)
        assert len(x) == len(y) == len(ldx) == len(ldy) == n
        abi.check(abi.lib().kg_field2d_set_object_locations(
            self._h, n, abi.ptr(ids), abi.ptr(x), abi.ptr(y), abi.ptr(ldx), abi.ptr(ldy)))

    def set_object_locations_dev(self, n, id_ptr, x_ptr, y_ptr, dx_ptr, dy_ptr):
        """Same with device pointers (e.g. torch tensors' data_ptr())."""
        abi.check(abi.lib().kg_field2d_set_object_locations_dev(self._h, n, id_ptr, x_ptr, y_ptr,
                                                                dx_ptr, dy_ptr))

    def remove_object_location(self, object, loc):
        """field_2d.rs:885-898"""
        oid, _, _ = _unpack(object)
        abi.check(abi.lib().kg_field2d_remove_object_location(self._h, oid, loc[0], loc[1]))

    def lazy_update(self):
        """Field::lazy_update  field_2d.rs:905-921: swap + cell-list rebuild."""
        abi.check(abi.lib().kg_field2d_lazy_update(self._h))

    def update(self):
        abi.check(abi.lib().kg_field2d_update(self._h))

    def set_order(self, canonical):
        abi.check(abi.lib().kg_field2d_set_order(
            self._h, abi.KG_ORDER_CANONICAL if canonical else abi.KG_ORDER_ANY))

    def set_kernel_variant(self, variant):
        """abi.KG_K4_AUTO / KG_K4_GENERIC / KG_K4_FAST_SCALAR / KG_K4_PACKED_BY_ID (same bits)"""
        abi.check(abi.lib().kg_field2d_set_kernel_variant(self._h, int(variant)))

    def sync(self):
        abi.check(abi.lib().kg_field2d_sync(self._h))

    # ------------------------------------------------------------------ read side
    @property
    def nagents(self):
        out = abi.u64()
        abi.check(abi.lib().kg_field2d_nagents(self._h, C.byref(out)))
        return out.value

    def num_objects(self, unbuffered=False):
        out = abi.u64()
        abi.check(abi.lib().kg_field2d_num_objects(self._h, int(unbuffered), C.byref(out)))
        return out.value

    def _neighbors(self, locs, dist, mode):
        locs = np.asarray(locs, dtype=np.float32).reshape(-1, 2)
        qx, qy = abi.as_f32(locs[:, 0]), abi.as_f32(locs[:, 1])
        nq = len(qx)
        offs = np.zeros(nq + 1, np.uint64)
        cap = max(1024, 64 * nq)
        while True:
            ids = np.zeros(cap, np.uint32)
            total = abi.u64()
            rc = abi.lib().kg_field2d_neighbors(self._h, nq, abi.ptr(qx), abi.ptr(qy), dist, mode,
                                                abi.ptr(offs), abi.ptr(ids), cap, C.byref(total))
            if rc == abi.KG_E_CAPACITY and total.value > cap:
                cap = total.value
                continue
            abi.check(rc)
            return offs.astype(np.int64), ids[: total.value]

    def get_neighbors_within_distance(self, loc, dist):
        """field_2d.rs:386-440 — ids of the neighbours: cells in the reference's order (x outer, y
        inner); inside a cell ascending id with set_order(canonical=True), otherwise the order the
        rebuild's atomics left (the reference's own bag order is its scheduler's push order)."""
        return self._neighbors([loc], dist, abi.KG_QUERY_EXACT)[1]

    def get_neighbors_within_relax_distance(self, loc, dist):
        """field_2d.rs:472-516 (same ordering note)"""
        return self._neighbors([loc], dist, abi.KG_QUERY_RELAX)[1]

    def neighbors_agents(self, locs, dist, exact):
        """The reference's queries return Vec<O>: (offsets, dict(id, x, y, ldx, ldy)) — the neighbours
        themselves for many query points, one boundary crossing."""
        locs = np.asarray(locs, dtype=np.float32).reshape(-1, 2)
        qx, qy = abi.as_f32(locs[:, 0]), abi.as_f32(locs[:, 1])
        nq = len(qx)
        offs = np.zeros(nq + 1, np.uint64)
        cap = max(1024, 64 * nq)
        mode = abi.KG_QUERY_EXACT if exact else abi.KG_QUERY_RELAX
        while True:
            ids = np.zeros(cap, np.uint32)
            a = [np.zeros(cap, np.float32) for _ in range(4)]
            total = abi.u64()
            rc = abi.lib().kg_field2d_neighbors_agents(self._h, nq, abi.ptr(qx), abi.ptr(qy), dist, mode,
                                                       abi.ptr(offs), abi.ptr(ids), *[abi.ptr(v) for v in a], cap,
                                                       C.byref(total))
            if rc == abi.KG_E_CAPACITY and total.value > cap:
                cap = total.value
                continue
            abi.check(rc)
            n = total.value
            return offs.astype(np.int64), dict(id=ids[:n], x=a[0][:n], y=a[1][:n], ldx=a[2][:n], ldy=a[3][:n])

    def neighbors_batch(self, locs, dist, exact):
        """(offsets, ids) CSR for many query points in one launch."""
        return self._neighbors(locs, dist, abi.KG_QUERY_EXACT if exact else abi.KG_QUERY_RELAX)

    def get_objects(self, loc, unbuffered=False):
        """field_2d.rs:546-560 (and get_objects_unbuffered :562-575)."""
        cap = 4096
        while True:
            ids = np.zeros(cap, np.uint32)
            n = abi.u64()
            rc = abi.lib().kg_field2d_get_objects(self._h, int(unbuffered), loc[0], loc[1], cap,
                                                  abi.ptr(ids), C.byref(n))
            if rc == abi.KG_E_CAPACITY and n.value > cap:
                cap = n.value
                continue
            abi.check(rc)
            return ids[: n.value]

    def get_objects_unbuffered(self, loc):
        return self.get_objects(loc, True)

    def num_objects_at_location(self, loc):
        """field_2d.rs:806-811"""
        x, y, out = abi.as_f32([loc[0]]), abi.as_f32([loc[1]]), np.zeros(1, np.uint32)
        abi.check(abi.lib().kg_field2d_num_objects_at_locations(self._h, 1, abi.ptr(x), abi.ptr(y),
                                                                abi.ptr(out)))
        return int(out[0])

    def num_empty_bags(self):
        out = abi.u64()
        abi.check(abi.lib().kg_field2d_num_empty_bags(self._h, C.byref(out)))
        return out.value

    def get_empty_bags(self):
        """field_2d.rs:718-730: cell origins (not_discretize) of the empty read bags."""
        counts = self.cell_counts()
        idx = np.nonzero(counts == 0)[0]
        d = np.float32(self.discretization)
        return [Real2D(float(np.float32(i // self.dh) * d), float(np.float32(i % self.dh) * d))
                for i in idx]

    def get_random_empty_bag(self, rng=None):
        """field_2d.rs:764-772: one of get_empty_bags() drawn uniformly, None when every bag is
        occupied.  `rng` is anything with randrange (default: the `random` module, OS-seeded like
        rand::rng())."""
        bags = self.get_empty_bags()
        if not bags:
            return None
        return bags[(rng or random).randrange(len(bags))]

    def cell_counts(self, unbuffered=False):
        out = np.zeros(self.dw * self.dh, np.uint32)
        abi.check(abi.lib().kg_field2d_cell_counts(self._h, int(unbuffered), len(out), abi.ptr(out)))
        return out

    def download(self, unbuffered=False, with_cells=True):
        """dict(id,x,y,ldx,ldy,cell): read buffer in iter_objects order (field_2d.rs:594-626)."""
        n = self.num_objects(unbuffered)
        a = dict(id=np.zeros(n, np.uint32), x=np.zeros(n, np.float32), y=np.zeros(n, np.float32),
                 ldx=np.zeros(n, np.float32), ldy=np.zeros(n, np.float32))
        cell = np.zeros(n, np.int32) if with_cells else None
        got = abi.u64()
        abi.check(abi.lib().kg_field2d_download(self._h, int(unbuffered), n, abi.ptr(a["id"]),
                                                abi.ptr(a["x"]), abi.ptr(a["y"]), abi.ptr(a["ldx"]),
                                                abi.ptr(a["ldy"]), abi.ptr(cell), C.byref(got)))
        if with_cells:
            a["cell"] = cell
        return a

    def iter_objects(self, closure, unbuffered=False):
        """field_2d.rs:594-660: closure(cell_origin: Real2D, object_id) in x-outer/y-inner order.
        (The write buffer is an append log on the device; it is presented cell-major here.)"""
        a = self.download(unbuffered)
        order = np.argsort(a["cell"], kind="stable") if unbuffered else np.arange(len(a["id"]))
        d = np.float32(self.discretization)
        for k in order:
            c = int(a["cell"][k])
            closure(Real2D(float(np.float32(c // self.dh) * d), float(np.float32(c % self.dh) * d)),
                    int(a["id"][k]))

    # ------------------------------------------------------------------ fused per-agent step
    def step_boids(self, params):
        """All agents' Bird::step (tests/model/flockers/bird.rs:39-155) in one launch."""
        abi.check(abi.lib().kg_field2d_step_boids(self._h, C.byref(params)))

    def step_boids_life(self, params, life):
        """Every agent's step + Agent::is_stopped (agent.rs:18), then the births of State::after_step
        (krabgpu.h KgLifeRule); returns (stopped, born) of this step.  Follow with lazy_update()."""
        ns, nb = abi.u64(), abi.u64()
        abi.check(abi.lib().kg_field2d_step_boids_life(self._h, C.byref(params), C.byref(life), C.byref(ns),
                                                       C.byref(nb)))
        return ns.value, nb.value

    def set_next_id(self, next_id):
        abi.check(abi.lib().kg_field2d_set_next_id(self._h, int(next_id)))

    def reduce(self):
        """Sums over the read buffer computed on the device (kg_field2d_reduce): sum_x, sum_y,
        sum_ldx, sum_ldy, sum_speed, sum_xx, sum_yy (f64, deterministic) and n."""
        out = np.zeros(8, np.float64)
        abi.check(abi.lib().kg_field2d_reduce(self._h, abi.ptr(out)))
        keys = ("sum_x", "sum_y", "sum_ldx", "sum_ldy", "sum_speed", "sum_xx", "sum_yy")
        red = {k: float(out[i]) for i, k in enumerate(keys)}
        red["n"] = self.num_objects()
        return red

    def run_boids(self, params, nsteps):
        abi.check(abi.lib().kg_field2d_run_boids(self._h, C.byref(params), nsteps))

    def step_custom(self, pair, finish, consts=(), radius=10.0, exact=False, seed=0, step=0, may_stop=False):
        """A model's own Agent::step for every agent: `pair` / `finish` are the step's body as CUDA C statements
        (include/krabgpu.h KgCustomStep), compiled at run time.  Reads the read buffer, pushes every agent's new
        copy into the write buffer; `lazy_update()` makes it current."""
        cs = abi.KgCustomStep()
        cs.pair, cs.finish = pair.encode(), finish.encode()
        for k, v in enumerate(consts):
            cs.consts[k] = float(v)
        cs.nconsts, cs.radius, cs.exact_query, cs.may_stop = len(consts), float(radius), int(bool(exact)), int(bool(may_stop))
        cs.seed, cs.step = int(seed), int(step)
        abi.check(abi.lib().kg_field2d_step_custom(self._h, C.byref(cs)))

    def run_boids_series(self, params, nsteps, every=1):
        """nsteps steps on the device; after every `every`-th step the reductions of `reduce()` are recorded
        in device memory and all rows come back with one copy at the end: array [nsteps // every, 8]
        (columns as in `reduce`, last one unused).  The data a `plot!` series is made of."""
        rows = nsteps // every
        out = np.zeros((max(rows, 1), 8), np.float64)
        abi.check(abi.lib().kg_field2d_run_boids_series(self._h, C.byref(params), nsteps, every, abi.ptr(out),
                                                        len(out)))
        return out[:rows]

    def init_flockers(self, n, seed):
        abi.check(abi.lib().kg_field2d_init_flockers(self._h, n, seed))

    def step_boids_host(self, params, inp, out):
        """e2e step with host SoA dicts (id,x,y,ldx,ldy), ideally pinned."""
        n = len(inp["id"])
        abi.check(abi.lib().kg_field2d_step_boids_host(
            self._h, C.byref(params), n, abi.ptr(inp["id"]), abi.ptr(inp["x"]), abi.ptr(inp["y"]),
            abi.ptr(inp["ldx"]), abi.ptr(inp["ldy"]), abi.ptr(out["id"]), abi.ptr(out["x"]),
            abi.ptr(out["y"]), abi.ptr(out["ldx"]), abi.ptr(out["ldy"])))

    def step_boids_host_ordered(self, params, inp, out):
        """e2e step for a host that finds its agents by POSITION: inp = dict(x, y, ldx, ldy[, id]); the result
        of input agent i lands at out[...][i] (out may be inp).  Without "id", agent i has id i."""
        n = len(inp["x"])
        ids = inp.get("id")
        abi.check(abi.lib().kg_field2d_step_boids_host_ordered(
            self._h, C.byref(params), n, abi.ptr(ids) if ids is not None else None, abi.ptr(inp["x"]),
            abi.ptr(inp["y"]), abi.ptr(inp["ldx"]), abi.ptr(inp["ldy"]), abi.ptr(out["x"]), abi.ptr(out["y"]),
            abi.ptr(out["ldx"]), abi.ptr(out["ldy"])))

    def l2_flush(self, nbytes=256 << 20):
        abi.check(abi.lib().kg_field2d_l2_flush(self._h, nbytes))

    def run_boids_timed(self, params, nsteps, flush_bytes=0):
        """run_boids with per-step CUDA-event timing; returns the summed device milliseconds."""
        ms = C.c_double()
        abi.check(abi.lib().kg_field2d_run_boids_timed(self._h, C.byref(params), nsteps, flush_bytes,
                                                       C.byref(ms)))
        return ms.value

    def timer_start(self):
        abi.check(abi.lib().kg_field2d_timer_start(self._h))

    def timer_stop(self):
        """milliseconds of device time since timer_start (CUDA events on the handle's stream)"""
        ms = C.c_double()
        abi.check(abi.lib().kg_field2d_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def profile(self, enable=True):
        abi.check(abi.lib().kg_field2d_profile(self._h, int(enable)))

    def profile_read(self, reset=True):
        ms = (C.c_double * 8)()
        ln = (abi.u64 * 8)()
        abi.check(abi.lib().kg_field2d_profile_read(self._h, ms, ln, int(reset)))
        return {k: (ms[i], ln[i]) for i, k in enumerate(abi.KERNEL_KINDS)}


def _unpack(obj):
    if hasattr(obj, "id"):
        ld = getattr(obj, "last_d", (0.0, 0.0))
        return int(obj.id), float(ld[0]), float(ld[1])
    if isinstance(obj, (tuple, list)):
        return int(obj[0]), float(obj[1]) if len(obj) > 1 else 0.0, float(obj[2]) if len(obj) > 2 else 0.0
    return int(obj), 0.0, 0.0
