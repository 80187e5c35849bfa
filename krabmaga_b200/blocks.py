"""2-D block decomposition of a Field2D world over the GPUs of one box (csrc/block.cu, SURVEY §8f-4;
precedent: the kd-tree blocks of src/engine/fields/kdtree_mpi.rs:211-238).  One process drives every
block; for worlds where strips of whole columns (strips.py) get too thin."""
import ctypes as C

import numpy as np

from . import _abi as abi


class BlockWorld:
    def __init__(self, w, h, discretization, radius, nbx, nby, devices, capacity, canonical_order=False,
                 slack=2.0, xcap=None):
        """`capacity` = agents of the whole world; every block is sized for its share x `slack` plus its halo
        ring.  `devices[k]` hosts block k = bx * nby + by (round-robin if shorter)."""
        self.w, self.h, self.nbx, self.nby = float(w), float(h), int(nbx), int(nby)
        self._blocks = []
        n = self.nbx * self.nby
        per = int(np.ceil(capacity / n * slack)) + 4096
        self._xcap = int(xcap or max(4096, per // 2))
        for k in range(n):
            hnd = abi.vp()
            abi.check(abi.lib().kg_block_create(w, h, discretization, 1, radius, k // self.nby, k % self.nby, self.nbx,
                                                self.nby, 2 * per, self._xcap, int(devices[k % len(devices)]),
                                                C.byref(hnd)))
            self._blocks.append(hnd)
            abi.check(abi.lib().kg_block_set_order(hnd, abi.KG_ORDER_CANONICAL if canonical_order else abi.KG_ORDER_ANY))
        self._arr = (abi.vp * n)(*self._blocks)
        self._cap = 2 * per

    def close(self):
        for hnd in getattr(self, "_blocks", []):
            abi.lib().kg_block_destroy(hnd)
        self._blocks = []

    __del__ = close

    def cells(self, k):
        own, loc = np.zeros(4, np.int32), np.zeros(4, np.int32)
        abi.check(abi.lib().kg_block_cells(self._blocks[k], abi.ptr(own), abi.ptr(loc)))
        return tuple(int(v) for v in own), tuple(int(v) for v in loc)

    def upload(self, agents):
        """N x set_object_location + lazy_update: every block keeps the agents inside its window"""
        ids = abi.as_u32(agents["id"])
        if len(np.unique(ids)) != len(ids):
            raise ValueError("BlockWorld needs unique agent ids (self exclusion is by buffer index)")
        x, y = abi.as_f32(agents["x"]), abi.as_f32(agents["y"])
        dx, dy = abi.as_f32(agents["ldx"]), abi.as_f32(agents["ldy"])
        for hnd in self._blocks:
            abi.check(abi.lib().kg_block_upload(hnd, len(ids), abi.ptr(ids), abi.ptr(x), abi.ptr(y), abi.ptr(dx),
                                                abi.ptr(dy)))
            abi.check(abi.lib().kg_block_lazy_update(hnd))

    def step_boids(self, params):
        abi.check(abi.lib().kg_blocks_step(self._arr, len(self._blocks), C.byref(params)))

    def run_boids(self, params, nsteps):
        abi.check(abi.lib().kg_blocks_run(self._arr, len(self._blocks), C.byref(params), nsteps))

    def counts(self):
        """[(agents held incl. ghosts, local cells)] per block"""
        out = []
        for hnd in self._blocks:
            a, c = abi.u64(), abi.u64()
            abi.check(abi.lib().kg_block_counts(hnd, C.byref(a), C.byref(c)))
            out.append((a.value, c.value))
        return out

    def download(self, per_block=False):
        """the owned agents of every block (ghosts left out): dict of arrays, or a list of them"""
        parts = []
        for hnd in self._blocks:
            cap = self._cap
            ids = np.zeros(cap, np.uint32)
            f = [np.zeros(cap, np.float32) for _ in range(4)]
            n = abi.u64()
            abi.check(abi.lib().kg_block_download(hnd, cap, abi.ptr(ids), abi.ptr(f[0]), abi.ptr(f[1]), abi.ptr(f[2]),
                                                  abi.ptr(f[3]), C.byref(n)))
            m = n.value
            parts.append(dict(id=ids[:m], x=f[0][:m], y=f[1][:m], ldx=f[2][:m], ldy=f[3][:m]))
        if per_block:
            return parts
        return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
