"""CPU tests of the multi-GPU host logic: strip partition, ownership, capacities and the IPC-handle
exchange over torch.distributed (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest

from krabmaga_b200 import strips

DISC = float(np.float32(10.0) / np.float32(1.5))


def test_partition_covers_every_column_once():
    for w, n in ((400.0, 1), (400.0, 2), (4000.0, 8), (32000.0, 8), (32000.0, 3)):
        max_x, _, dw, _ = strips.grid_dims(w, w, DISC)
        parts = strips.partition(w, w, DISC, n)
        assert parts[0][0] == 0 and parts[-1][1] == dw        # padding column goes to the last rank
        for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
            assert a1 == b0 and a1 > a0
        assert sum(x1 - x0 for x0, x1 in parts) == dw


def test_grid_dims_follow_reference_f32_rounding():
    # 400 / 6.6666665f must give max_x = 60 (61 would be a different grid)
    assert strips.grid_dims(400.0, 400.0, DISC) == (60, 60, 61, 61)
    assert strips.grid_dims(4000.0, 4000.0, DISC)[0] == 600
    assert strips.grid_dims(32000.0, 32000.0, DISC)[0] == 4800
    assert strips.grid_dims(10.0, 10.0, 0.5) == (20, 20, 21, 21)


def test_owner_of_matches_partition():
    w, n = 4000.0, 8
    parts = strips.partition(w, w, DISC, n)
    rng = np.random.default_rng(0)
    x = np.concatenate([(rng.random(5000) * w).astype(np.float32), np.float32([0.0, w, 3999.9998])])
    own = strips.owner_of(x, w, w, DISC, n)
    cols = np.floor(x / np.float32(DISC)).astype(int)
    for c, r in zip(cols, own):
        assert parts[r][0] <= c < parts[r][1]
    assert own[-2] == n - 1   # x == w sits in the padding column, owned by the last rank


def test_default_capacities_are_sane():
    cap, hcap, mcap = strips.default_capacities(64_000_000, 32000.0, 32000.0, DISC, 10.0, 8)
    assert 8_000_000 < cap < 16_000_000
    assert hcap > 4 * 64_000_000 // 4800 and mcap > 64_000_000 // 4800


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bytes([rank]) * strips.KG_IPC_HANDLE_BYTES
    left, right = strips.exchange_handles(mine, rank, world, dist)
    # every rank also partitions the same seeded population and reports what it would upload
    rng = np.random.default_rng(1)
    x = (rng.random(10000) * 400.0).astype(np.float32)
    own = strips.owner_of(x, 400.0, 400.0, DISC, world)
    q.put((rank, left[0], right[0], int((own == rank).sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_handle_exchange_over_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(g[3] for g in got) == 10000                      # every agent has exactly one owner
    for rank, left, right, _ in got:
        assert left == (rank - 1) % world and right == (rank + 1) % world
