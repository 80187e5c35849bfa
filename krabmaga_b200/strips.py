"""Multi-GPU Field2D: x-strip decomposition with NVLink halo exchange and agent migration.

Host side of include/krabgpu.h's kg_strip_* group.  The reference's precedent is the MPI
`Kdtree` field (src/engine/fields/kdtree_mpi.rs): a block of the world per rank, halo regions of
width `distance`, a per-step exchange.  Here a rank is a GPU, the block is a strip of whole cell
columns (contiguous in the x-major cell order) and the exchange is peer stores issued
by the step's own kernels.

Two ways to run it:
  * one process per GPU (torchrun): `StripField2D` + `connect_ipc(strip, dist)`;
  * one process driving several devices (or several strips on one device, for tests): `StripWorld`.
"""
import ctypes as C

import numpy as np

from . import _abi as abi

KG_IPC_HANDLE_BYTES = 64


# ------------------------------------------------------------------------------ host-side geometry
def grid_dims(w, h, disc):
    """(max_x, max_y, dw, dh) with the f32 operation order of field_2d.rs:317-318 / :487-488."""
    w, h, d = np.float32(w), np.float32(h), np.float32(disc)
    max_x = int(np.ceil(w / d))
    max_y = int(np.ceil(h / d))
    return max_x, max_y, max_x + 1, max_y + 1


def partition(w, h, disc, nranks):
    """Owned global cell-column range [x0, x1) of every rank; the last rank also owns the padding
    column max_x.  Mirrors col_range() in csrc/strip.cu."""
    max_x, _, dw, _ = grid_dims(w, h, disc)
    out = []
    for r in range(nranks):
        x0 = r * max_x // nranks
        x1 = dw if r == nranks - 1 else (r + 1) * max_x // nranks
        out.append((x0, x1))
    return out


def owner_of(x, w, h, disc, nranks):
    """Rank owning each x coordinate: discretize (field_2d.rs:328-339) then look up the strip."""
    cols = np.floor(np.asarray(x, np.float32) / np.float32(disc)).astype(np.int64)
    bounds = np.array([p[0] for p in partition(w, h, disc, nranks)] + [1 << 62])
    return np.searchsorted(bounds, cols, side="right") - 1


def default_capacities(n_global, w, h, disc, radius, nranks, slack=1.5):
    """(capacity, halo_capacity, migrate_capacity) for a roughly uniform population."""
    max_x, _, dw, dh = grid_dims(w, h, disc)
    dd = int(np.floor(np.float32(radius) / np.float32(disc)))
    per_col = n_global / max(max_x, 1)
    widest = max(x1 - x0 for x0, x1 in partition(w, h, disc, nranks))
    capacity = int(per_col * widest * slack) + 1024
    halo = int(per_col * max(dd, 1) * 4 * slack) + 1024
    migrate = int(per_col * 2 * slack) + 1024  # |step| <= jump < one column
    return min(capacity, max(n_global, 1024) + 1024), halo, migrate


# ------------------------------------------------------------------------------ one strip
class StripField2D:
    """One rank's strip of a multi-GPU toroidal Field2D (relaxed query)."""

    def __init__(self, w, h, disc, radius, rank, nranks, capacity, halo_capacity, migrate_capacity,
                 device=0, toroidal=True):
        self._h = abi.vp()
        abi.check(abi.lib().kg_strip_create(w, h, disc, int(toroidal), radius, rank, nranks, capacity,
                                            halo_capacity, migrate_capacity, device, C.byref(self._h)))
        self.rank, self.nranks, self.device = rank, nranks, device
        self.width, self.height, self.discretization, self.radius = w, h, disc, radius
        self.capacity = capacity
        v = [abi.i32() for _ in range(5)]
        abi.check(abi.lib().kg_strip_columns(self._h, *[C.byref(a) for a in v]))
        self.own_x0, self.own_x1, self.halo_l, self.halo_r, self.dh = [a.value for a in v]

    def close(self):
        if getattr(self, "_h", None):
            abi.lib().kg_strip_destroy(self._h)
            self._h = None

    __del__ = close

    def set_order(self, canonical):
        abi.check(abi.lib().kg_strip_set_order(
            self._h, abi.KG_ORDER_CANONICAL if canonical else abi.KG_ORDER_ANY))

    def ipc_export(self):
        buf = (C.c_ubyte * KG_IPC_HANDLE_BYTES)()
        abi.check(abi.lib().kg_strip_ipc_export(self._h, buf))
        return bytes(buf)

    def connect_ipc(self, left_handle, right_handle):
        lh = (C.c_ubyte * KG_IPC_HANDLE_BYTES).from_buffer_copy(left_handle)
        rh = (C.c_ubyte * KG_IPC_HANDLE_BYTES).from_buffer_copy(right_handle)
        abi.check(abi.lib().kg_strip_connect_ipc(self._h, lh, rh))

    def connect_local(self, left, right):
        abi.check(abi.lib().kg_strip_connect_local(self._h, left._h, right._h))

    def init_flockers(self, n_global, seed):
        abi.check(abi.lib().kg_strip_init_flockers(self._h, n_global, seed))

    def upload(self, ids, x, y, ldx, ldy):
        ids = abi.as_u32(ids)
        a = [abi.as_f32(v) for v in (x, y, ldx, ldy)]
        abi.check(abi.lib().kg_strip_upload(self._h, len(ids), abi.ptr(ids), *[abi.ptr(v) for v in a]))

    def clear(self):
        abi.check(abi.lib().kg_strip_clear(self._h))

    def prepare(self):
        abi.check(abi.lib().kg_strip_prepare(self._h))

    def step_boids(self, params):
        abi.check(abi.lib().kg_strip_step_boids(self._h, C.byref(params)))

    def run_boids(self, params, nsteps):
        abi.check(abi.lib().kg_strip_run_boids(self._h, C.byref(params), nsteps))

    def run_boids_timed(self, params, nsteps, flush_bytes=0):
        ms = C.c_double()
        abi.check(abi.lib().kg_strip_run_boids_timed(self._h, C.byref(params), nsteps, flush_bytes,
                                                     C.byref(ms)))
        return ms.value

    def sync(self):
        abi.check(abi.lib().kg_strip_sync(self._h))

    def stats(self):
        v = [abi.u64() for _ in range(6)]
        abi.check(abi.lib().kg_strip_stats(self._h, *[C.byref(a) for a in v]))
        keys = ("n_owned", "migrants_in", "migrants_out", "halo_left", "halo_right", "launches")
        return dict(zip(keys, (a.value for a in v)))

    def download(self, out=None):
        """Owned agents; `out` may be a dict of preallocated (e.g. pinned) arrays of >= n_owned
        entries, in which case views of length n_owned are returned."""
        n = self.stats()["n_owned"]
        if out is None:
            a = dict(id=np.zeros(n, np.uint32), x=np.zeros(n, np.float32), y=np.zeros(n, np.float32),
                     ldx=np.zeros(n, np.float32), ldy=np.zeros(n, np.float32))
        else:
            a = {k: v[:n] for k, v in out.items()}
        got = abi.u64()
        abi.check(abi.lib().kg_strip_download(self._h, n, abi.ptr(a["id"]), abi.ptr(a["x"]),
                                              abi.ptr(a["y"]), abi.ptr(a["ldx"]), abi.ptr(a["ldy"]),
                                              C.byref(got)))
        return a

    def timer_start(self):
        abi.check(abi.lib().kg_strip_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double()
        abi.check(abi.lib().kg_strip_timer_stop(self._h, C.byref(ms)))
        return ms.value


def exchange_handles(my_handle, rank, nranks, dist=None):
    """All-gather the 64-byte IPC handles and return (left_handle, right_handle) of the ring
    neighbours.  `dist` is torch.distributed (any backend); None means a single rank."""
    if nranks == 1 or dist is None:
        return my_handle, my_handle
    handles = [None] * nranks
    dist.all_gather_object(handles, my_handle)
    return handles[(rank - 1) % nranks], handles[(rank + 1) % nranks]


def connect_ipc(strip, dist=None):
    """Wire one-process-per-GPU strips together through torch.distributed."""
    left, right = exchange_handles(strip.ipc_export(), strip.rank, strip.nranks, dist)
    if strip.nranks > 1:
        strip.connect_ipc(left, right)
    if dist is not None and strip.nranks > 1:
        dist.barrier()


# ------------------------------------------------------------------------------ several strips, one process
class StripWorld:
    """All strips of a world driven from one process (the shape a single-process Rust host takes:
    `kg_world` in SURVEY §8b).  `devices[r]` is the CUDA device of strip r; repeating a device is
    allowed (used by the single-GPU tests)."""

    def __init__(self, w, h, disc, radius, devices, n_global, canonical_order=False, slack=1.5):
        self.w, self.h, self.disc, self.radius = w, h, disc, radius
        self.nranks = len(devices)
        cap, hcap, mcap = default_capacities(n_global, w, h, disc, radius, self.nranks, slack)
        self.strips = [StripField2D(w, h, disc, radius, r, self.nranks, cap, hcap, mcap, device=d)
                       for r, d in enumerate(devices)]
        for s in self.strips:
            s.set_order(canonical_order)
        if self.nranks > 1:
            for r, s in enumerate(self.strips):
                s.connect_local(self.strips[(r - 1) % self.nranks], self.strips[(r + 1) % self.nranks])

    def close(self):
        for s in self.strips:
            s.close()

    def init_flockers(self, n_global, seed):
        for s in self.strips:
            s.init_flockers(n_global, seed)
        self.prepare()

    def upload(self, agents):
        """Scatter a host population (dict id,x,y,ldx,ldy) to the owning strips."""
        own = owner_of(agents["x"], self.w, self.h, self.disc, self.nranks)
        for r, s in enumerate(self.strips):
            m = own == r
            s.upload(agents["id"][m], agents["x"][m], agents["y"][m], agents["ldx"][m], agents["ldy"][m])
        self.prepare()

    def prepare(self):
        for s in self.strips:
            s.prepare()

    def run_boids(self, params, nsteps):
        """Step s is issued for every strip before step s+1 for any (see kg_strip_step_boids)."""
        step0 = params.step
        for i in range(nsteps):
            params.step = step0 + i
            for s in self.strips:
                s.step_boids(params)
        params.step = step0
        self.sync()

    def sync(self):
        for s in self.strips:
            s.sync()

    def download(self):
        parts = [s.download() for s in self.strips]
        return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}

    def stats(self):
        return [s.stats() for s in self.strips]
