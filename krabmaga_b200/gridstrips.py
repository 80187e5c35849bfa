"""Multi-GPU DenseNumberGrid2D<u8> (include/krabgpu.h: kg_gridstrip_*): strips of whole x rows,
one per GPU, halo rows exchanged inside the stencil kernel by peer stores over NVLink.

StripDenseNumberGrid2D  one strip (one per process under torchrun)
GridStripWorld          all strips of a grid driven from one process
row_range / connect_ipc host-side partition and wiring helpers (pure Python, testable on CPU)
"""
import ctypes as C

import numpy as np

from . import _abi as abi

KG_IPC_HANDLE_BYTES = 64


def row_range(width, rank, nranks):
    """rows [x0, x1) owned by `rank` — the same integer formula as kg_gridstrip_create"""
    return rank * width // nranks, (rank + 1) * width // nranks


class StripDenseNumberGrid2D:
    def __init__(self, width, height, rank, nranks, device=0):
        self._h = abi.vp()
        abi.check(abi.lib().kg_gridstrip_create(width, height, rank, nranks, device, C.byref(self._h)))
        self.width, self.height, self.rank, self.nranks, self.device = width, height, rank, nranks, device
        x0, x1 = abi.i32(), abi.i32()
        abi.check(abi.lib().kg_gridstrip_rows(self._h, x0, x1))
        self.x0, self.x1 = x0.value, x1.value
        assert (self.x0, self.x1) == row_range(width, rank, nranks)

    def close(self):
        if getattr(self, "_h", None):
            abi.lib().kg_gridstrip_destroy(self._h)
            self._h = None

    __del__ = close

    def ipc_export(self):
        buf = (C.c_ubyte * KG_IPC_HANDLE_BYTES)()
        abi.check(abi.lib().kg_gridstrip_ipc_export(self._h, buf))
        return bytes(buf)

    def connect_ipc(self, left_handle, right_handle):
        def h(b):
            return None if b is None else (C.c_ubyte * KG_IPC_HANDLE_BYTES).from_buffer_copy(b)
        abi.check(abi.lib().kg_gridstrip_connect_ipc(self._h, h(left_handle), h(right_handle)))

    def connect_local(self, left, right):
        abi.check(abi.lib().kg_gridstrip_connect_local(self._h, left._h if left else None,
                                                       right._h if right else None))

    def init_forest_fire(self, density, seed):
        abi.check(abi.lib().kg_gridstrip_init_forest_fire(self._h, density, seed))

    def upload(self, own_rows):
        a = np.ascontiguousarray(own_rows, np.uint8).reshape(-1)
        assert a.size == (self.x1 - self.x0) * self.height
        abi.check(abi.lib().kg_gridstrip_upload(self._h, abi.ptr(a)))

    def download(self, out=None):
        """the strip's own rows; `out`: a flat uint8 buffer to fill (page-locked for full PCIe rate)"""
        n = (self.x1 - self.x0) * self.height
        if out is None:
            out = np.empty(n, np.uint8)
        assert out.dtype == np.uint8 and out.size == n and out.flags["C_CONTIGUOUS"]
        abi.check(abi.lib().kg_gridstrip_download(self._h, abi.ptr(out)))
        return out.reshape(self.x1 - self.x0, self.height)

    def prepare(self):
        abi.check(abi.lib().kg_gridstrip_prepare(self._h))

    def run_stencil(self, nsteps, rule=abi.KG_RULE_FOREST_FIRE):
        abi.check(abi.lib().kg_gridstrip_run_stencil(self._h, rule, nsteps))

    def run_stencil_timed(self, nsteps, rule=abi.KG_RULE_FOREST_FIRE):
        ms = C.c_double()
        abi.check(abi.lib().kg_gridstrip_run_stencil_timed(self._h, rule, nsteps, C.byref(ms)))
        return ms.value

    def sync(self):
        abi.check(abi.lib().kg_gridstrip_sync(self._h))


def line_neighbours(handles, rank):
    """(left, right) entries of `handles` for a line topology; None at the ends"""
    n = len(handles)
    return (handles[rank - 1] if rank > 0 else None, handles[rank + 1] if rank < n - 1 else None)


def connect_ipc(strip, dist=None):
    """Wire one-process-per-GPU strips together through torch.distributed (any backend)."""
    if strip.nranks == 1 or dist is None:
        return
    handles = [None] * strip.nranks
    dist.all_gather_object(handles, strip.ipc_export())
    left, right = line_neighbours(handles, strip.rank)
    strip.connect_ipc(left, right)
    dist.barrier()


def pass_plan(width, height, nranks, nsteps):
    """[(steps, rows per tile), ...] of kg_gridstrip_run_stencil(nsteps) — host logic, no device needed"""
    n = abi.u64()
    cap = max(1, int(nsteps))
    steps = np.zeros(cap, np.int32)
    rows = np.zeros(cap, np.int32)
    abi.check(abi.lib().kg_gridstrip_pass_plan(width, height, nranks, nsteps, abi.ptr(steps), abi.ptr(rows), cap,
                                               C.byref(n)))
    return [(int(steps[k]), int(rows[k])) for k in range(n.value)]


class GridStripWorld:
    """All strips of one grid in one process; `devices[r]` may repeat (single-GPU tests)."""
    CHUNK = 8  # steps enqueued per strip before moving on: a strip's kernels wait for its neighbours'

    def __init__(self, width, height, devices):
        self.width, self.height, self.nranks = width, height, len(devices)
        self.strips = [StripDenseNumberGrid2D(width, height, r, self.nranks, d)
                       for r, d in enumerate(devices)]
        for r, s in enumerate(self.strips):
            left, right = line_neighbours(self.strips, r)
            s.connect_local(left, right)

    def close(self):
        for s in self.strips:
            s.close()

    def init_forest_fire(self, density, seed):
        for s in self.strips:
            s.init_forest_fire(density, seed)
        self.prepare()

    def upload(self, cells):
        cells = np.ascontiguousarray(cells, np.uint8).reshape(self.width, self.height)
        for s in self.strips:
            s.upload(cells[s.x0:s.x1])
        self.prepare()

    def prepare(self):
        for s in self.strips:
            s.prepare()

    def run_stencil(self, nsteps):
        done = 0
        while done < nsteps:
            k = min(self.CHUNK, nsteps - done)
            for s in self.strips:
                s.run_stencil(k)
            done += k

    def download(self):
        return np.concatenate([s.download() for s in self.strips], axis=0)
