"""Host logic of the 2-D block decomposition (csrc/block.cu): the partition rule — every column owned by exactly
one part, the last part also the padding column, halo rings clipped to the world — checked without a GPU."""
import ctypes as C

import numpy as np
import pytest

from krabmaga_b200 import _abi as abi


def part(b, parts, maxc, dd, c=0):
    out, owner = np.zeros(4, np.int32), abi.i32()
    abi.check(abi.lib().kg_block_partition(b, parts, maxc, dd, c, abi.ptr(out), C.byref(owner)))
    return tuple(int(v) for v in out), owner.value


@pytest.mark.parametrize("parts,maxc,dd", [(1, 60, 1), (2, 60, 1), (3, 91, 1), (4, 600, 1), (8, 4800, 1), (3, 200, 3),
                                           (7, 61, 2), (5, 5, 0)])
def test_partition_covers_every_column_once(parts, maxc, dd):
    seen = np.zeros(maxc + 1, np.int32)
    for b in range(parts):
        (x0, x1, lo, hi), _ = part(b, parts, maxc, dd)
        assert 0 <= x0 < x1 <= maxc + 1
        seen[x0:x1] += 1
        assert lo == max(x0 - dd, 0) and hi == min(x1 + dd, maxc + 1)      # halo ring clipped to the world
        assert x0 == b * maxc // parts                                        # the strips' rule (strips.partition)
    assert (seen == 1).all()
    assert part(parts - 1, parts, maxc, dd)[0][1] == maxc + 1                # the padding column
    for c in range(maxc + 1):
        _, owner = part(0, parts, maxc, dd, c)
        (x0, x1, _, _), _ = part(owner, parts, maxc, dd)
        assert x0 <= c < x1


def test_partition_rejects_nonsense():
    out, owner = np.zeros(4, np.int32), abi.i32()
    assert abi.lib().kg_block_partition(3, 3, 10, 1, 0, abi.ptr(out), C.byref(owner)) == abi.KG_E_INVALID
    assert abi.lib().kg_block_partition(0, 0, 10, 1, 0, abi.ptr(out), C.byref(owner)) == abi.KG_E_INVALID
