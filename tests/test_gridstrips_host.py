"""Host-side logic of the grid strips (no device): partition, line topology, and the handle
exchange over torch.distributed with the gloo backend at world_size 2."""
import os
import subprocess
import sys

from krabmaga_b200.gridstrips import line_neighbours, pass_plan, row_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_ranges_tile_the_grid():
    for width, G in ((32768, 8), (130, 3), (9, 8), (5, 5)):
        edges = [row_range(width, r, G) for r in range(G)]
        assert edges[0][0] == 0 and edges[-1][1] == width
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        assert all(x1 > x0 for x0, x1 in edges)


def test_line_topology_has_open_ends():
    h = ["a", "b", "c"]
    assert line_neighbours(h, 0) == (None, "b")
    assert line_neighbours(h, 1) == ("a", "c")
    assert line_neighbours(h, 2) == ("b", None)


def test_pass_plan_of_fused_stencil_steps():
    """run_stencil's host logic: passes of 8 / 4 / 2 / 1 steps that add up, never more steps per pass than
    the smallest strip has rows (its neighbours need that many halo rows from it), and row tiles whose last
    one keeps the eight rows a pass hands over (or a single tile)"""
    for width, height, G, n in ((32768, 32768, 8, 1000), (32768, 32768, 1, 203), (130, 960, 2, 37), (12, 32, 6, 9),
                                (9, 16, 8, 5), (64, 2064, 4, 13), (131, 496, 2, 1), (4096, 4096, 3, 0)):
        plan = pass_plan(width, height, G, n)
        assert sum(t for t, _ in plan) == n
        own_min = width // G
        ts = [t for t, _ in plan]
        assert ts == sorted(ts, reverse=True)                       # 8s, then 4, 2, 1
        for t, rows in plan:
            assert t in (8, 4, 2, 1) and (t == 1 or t <= own_min)
            assert rows >= 8
            assert own_min <= rows or own_min % rows == 0 or own_min % rows >= 8
    assert [t for t, _ in pass_plan(32768, 32768, 8, 23)] == [8, 8, 4, 2, 1]
    assert [t for t, _ in pass_plan(12, 32, 6, 5)] == [2, 2, 1]      # two-row strips
    assert [t for t, _ in pass_plan(9, 16, 8, 3)] == [1, 1, 1]       # one-row strips never fuse


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from krabmaga_b200.gridstrips import line_neighbours
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
handles = [None] * world
dist.all_gather_object(handles, bytes([rank]) * 64)
left, right = line_neighbours(handles, rank)
assert left == (None if rank == 0 else bytes([rank - 1]) * 64)
assert right == (None if rank == world - 1 else bytes([rank + 1]) * 64)
dist.barrier()
dist.destroy_process_group()
sys.stdout.write(f"rank{rank}-ok\n")
'''


def test_handle_exchange_gloo_world_size_2(tmp_path):
    import socket
    with socket.socket() as sock:       # a free port, so parallel test sessions cannot collide
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank0-ok" in out.stdout and "rank1-ok" in out.stdout
