"""Batched Flockers replicas on one GPU — the state type the `explore` mirror instantiates.

`FlockerBatch` is R independent `Flocker` states (tests/model/flockers/state.rs) of the same world
and population, advanced together by one launch per phase (include/krabgpu.h: kg_batch_*).  It is
what `explore_parallel!` (src/explore/model_exploration.rs:354-423) fans out over rayon tasks, for
models whose `Agent::step` is the shipped boids kernel.
"""
import ctypes as C

import numpy as np

from . import _abi as abi


class FlockerBatch:
    def __init__(self, dim, initial_flockers, replicas, discretization, toroidal=True, params=None,
                 device=0, canonical_order=False):
        """`params`: one KgBoidsParams per replica (the swept inputs), or None for the fixture's
        constants with seed 42 + replica index."""
        self._h = abi.vp()
        abi.check(abi.lib().kg_batch_create(dim[0], dim[1], discretization, int(bool(toroidal)),
                                            replicas, initial_flockers, device, C.byref(self._h)))
        self.dim, self.replicas, self.initial_flockers = dim, int(replicas), int(initial_flockers)
        self.device, self.step = device, 0
        if canonical_order:
            abi.check(abi.lib().kg_batch_set_order(self._h, abi.KG_ORDER_CANONICAL))
        if params is not None:
            self.set_params(0, params)

    def close(self):
        if getattr(self, "_h", None):
            abi.lib().kg_batch_destroy(self._h)
            self._h = None

    __del__ = close

    def set_params(self, first, params):
        arr = (abi.KgBoidsParams * len(params))(*params)
        abi.check(abi.lib().kg_batch_set_params(self._h, first, len(params), arr))

    # ---- State::init / Schedule::step / State::update for every replica at once
    def init(self):
        """state.rs:41-56 per replica (Philox key = that replica's seed), then the first swap
        (schedule.rs:349-351)."""
        abi.check(abi.lib().kg_batch_init_flockers(self._h))
        abi.check(abi.lib().kg_batch_lazy_update(self._h))
        self.step = 0

    def upload(self, ids, x, y, ldx, ldy):
        """replica-major arrays of replicas * initial_flockers entries; replaces the population"""
        a = [abi.as_u32(ids)] + [abi.as_f32(v) for v in (x, y, ldx, ldy)]
        assert all(len(v) == self.replicas * self.initial_flockers for v in a)
        abi.check(abi.lib().kg_batch_upload(self._h, *[abi.ptr(v) for v in a]))
        abi.check(abi.lib().kg_batch_lazy_update(self._h))
        self.step = 0

    def step_once(self):
        abi.check(abi.lib().kg_batch_step_boids(self._h, self.step))
        abi.check(abi.lib().kg_batch_lazy_update(self._h))
        self.step += 1

    def run(self, nsteps):
        abi.check(abi.lib().kg_batch_run_boids(self._h, self.step, nsteps))
        self.step += nsteps

    def run_timed(self, nsteps, flush_bytes=0):
        ms = C.c_double()
        abi.check(abi.lib().kg_batch_run_boids_timed(self._h, self.step, nsteps, flush_bytes, C.byref(ms)))
        self.step += nsteps
        return ms.value

    def download(self, with_cells=False):
        """dict of [replicas, initial_flockers] arrays, each replica in its iter_objects order"""
        n = self.replicas * self.initial_flockers
        a = dict(id=np.zeros(n, np.uint32), x=np.zeros(n, np.float32), y=np.zeros(n, np.float32),
                 ldx=np.zeros(n, np.float32), ldy=np.zeros(n, np.float32))
        cell = np.zeros(n, np.int32) if with_cells else None
        abi.check(abi.lib().kg_batch_download(self._h, abi.ptr(a["id"]), abi.ptr(a["x"]), abi.ptr(a["y"]),
                                              abi.ptr(a["ldx"]), abi.ptr(a["ldy"]), abi.ptr(cell)))
        if with_cells:
            a["cell"] = cell
        return {k: v.reshape(self.replicas, self.initial_flockers) for k, v in a.items()}

    def reduce(self):
        """Per-replica sums over the read buffer, computed on the device (kg_batch_reduce): dict of
        [replicas] f64 arrays sum_x, sum_y, sum_ldx, sum_ldy, sum_speed, sum_xx, sum_yy plus n —
        what output columns are made of, for replicas x 64 bytes of download."""
        out = np.zeros((self.replicas, 8), np.float64)
        abi.check(abi.lib().kg_batch_reduce(self._h, abi.ptr(out)))
        keys = ("sum_x", "sum_y", "sum_ldx", "sum_ldy", "sum_speed", "sum_xx", "sum_yy")
        red = {k: out[:, i].copy() for i, k in enumerate(keys)}
        red["n"] = np.full(self.replicas, self.initial_flockers, np.int64)
        return red

    def sync(self):
        abi.check(abi.lib().kg_batch_sync(self._h))

    def timer_start(self):
        abi.check(abi.lib().kg_batch_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double()
        abi.check(abi.lib().kg_batch_timer_stop(self._h, C.byref(ms)))
        return ms.value
