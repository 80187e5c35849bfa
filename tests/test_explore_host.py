"""Host-side logic of the explore mirror (no device needed)."""
import pytest

from krabmaga_b200.explore import ExploreMode, build_configurations, deal_runs


def test_exaustive_order_matches_build_configurations_macro():
    """src/lib.rs:1724-1748: the first input varies slowest"""
    c = build_configurations({"cohesion": [1, 2], "avoidance": [10, 20, 30]}, ExploreMode.Exaustive)
    assert [(d["cohesion"], d["avoidance"]) for d in c] == \
        [(1, 10), (1, 20), (1, 30), (2, 10), (2, 20), (2, 30)]


def test_matched_zips_and_checks_lengths():
    c = build_configurations({"cohesion": [1, 2], "seed": [5, 6]}, ExploreMode.Matched)
    assert c == [{"cohesion": 1, "seed": 5}, {"cohesion": 2, "seed": 6}]
    with pytest.raises(ValueError):
        build_configurations({"cohesion": [1, 2], "seed": [5]}, ExploreMode.Matched)
    with pytest.raises(ValueError):
        build_configurations({"nope": [1]}, ExploreMode.Matched)


def test_runs_are_dealt_round_robin_and_batched():
    """run i -> device i % G (explore/mpi/model_exploration.rs:206), shares cut into batches"""
    plan = deal_runs(10, 3, 2)
    assert plan == [(0, [0, 3]), (0, [6, 9]), (1, [1, 4]), (1, [7]), (2, [2, 5]), (2, [8])]
    assert sorted(k for _, chunk in deal_runs(4096, 8, 512) for k in chunk) == list(range(4096))
    assert all(len(chunk) == 512 for _, chunk in deal_runs(4096, 8, 512))
    assert deal_runs(0, 4, 8) == []


def test_write_csv_header_then_one_record_per_row(tmp_path):
    """write_csv (src/lib.rs:1781-1792): `<name>.csv`, DataFrame::field_names as the header, every
    field through its string form"""
    import csv

    import numpy as np

    from krabmaga_b200.explore import field_names, write_csv
    rows = [dict(conf_num=0, conf_rep=1, cohesion=1.5, polarisation=np.float32(0.25), run_duration=0.5,
                 step_per_sec=400.0),
            dict(conf_num=1, conf_rep=0, cohesion=2.0, polarisation=np.float32(0.1), run_duration=0.25,
                 step_per_sec=800.0)]
    assert field_names(rows) == ["conf_num", "conf_rep", "cohesion", "polarisation", "run_duration",
                                 "step_per_sec"]
    path = write_csv(str(tmp_path / "explore_result"), rows)
    assert path.endswith("explore_result.csv")
    got = list(csv.reader(open(path)))
    assert got[0] == field_names(rows)
    assert got[1] == ["0", "1", "1.5", "0.25", "0.5", "400.0"]
    assert [float(v) for v in got[2]] == [1, 0, 2.0, float(np.float32(0.1)), 0.25, 800.0]
    # an empty dataframe still creates the file
    assert list(csv.reader(open(write_csv(str(tmp_path / "empty"), [])))) == [[]]
    with pytest.raises(ValueError):
        write_csv(str(tmp_path / "ragged"), [dict(a=1), dict(b=2)])


# ---- the sweep drivers above a stand-in for the device batch (the real one needs a GPU):
# what is tested here is the host logic — run order, dealing, batching, seeds, gathering.
class FakeBatch:
    made = []

    def __init__(self, dim, n, replicas, disc, toroidal, params, device=0, canonical_order=False):
        self.n, self.params, self.device = n, list(params), device
        assert len(self.params) == replicas
        FakeBatch.made.append((device, replicas))

    def init(self):
        pass

    def sync(self):
        pass

    def run(self, nstep):
        self.nstep = nstep

    def close(self):
        pass

    def reduce(self):
        d = self.download()
        import numpy as np
        return dict(sum_ldx=d["ldx"].astype(np.float64).sum(axis=1), sum_ldy=d["ldy"].astype(np.float64).sum(axis=1),
                    n=np.full(len(self.params), self.n))

    def download(self):
        import numpy as np
        # last_d = 0.7 * (cohesion, 0) / 4 for every agent, plus the seed in the y component
        ldx = np.stack([np.full(self.n, 0.7 * p.cohesion / 4, np.float32) for p in self.params])
        ldy = np.stack([np.full(self.n, 0.7 * (p.seed % 5) / 10, np.float32) for p in self.params])
        return dict(ldx=ldx, ldy=ldy)


def expected_polarisation(cohesion, seed):
    import numpy as np
    vx, vy = np.float32(0.7 * cohesion / 4), np.float32(0.7 * (seed % 5) / 10)
    return float(np.sqrt(vx * vx + vy * vy) / 0.7)


def test_explore_parallel_rows_dealing_and_seeds(monkeypatch):
    """explore_parallel! (model_exploration.rs:387-420): n_conf * rep_conf rows in run order, run i on
    device i % G, batches capped, every (seed, configuration, repetition) runs on its own stream"""
    import krabmaga_b200.explore as ex
    monkeypatch.setattr(ex, "FlockerBatch", FakeBatch)
    FakeBatch.made = []
    rows = ex.explore_parallel(7, 3, (100.0, 100.0), 50, 5.0, {"cohesion": [1.0, 2.0, 3.0], "seed": [10, 20, 30]},
                               mode=ExploreMode.Matched, devices=(0, 1), max_replicas_per_batch=2)
    assert [(r["conf_num"], r["conf_rep"]) for r in rows] == [(i, k) for i in range(3) for k in range(3)]
    assert [(r["cohesion"], r["seed"]) for r in rows[::3]] == [(1.0, 10), (2.0, 20), (3.0, 30)]
    for r in rows:
        eff = ex.run_seed(r["seed"], r["conf_num"], r["conf_rep"])
        assert r["effective_seed"] == eff
        assert r["polarisation"] == pytest.approx(expected_polarisation(r["cohesion"], eff))
        assert r["step_per_sec"] == pytest.approx(7 / r["run_duration"])
    # 9 runs over 2 devices: device 0 gets runs 0,2,4,6,8 (batches 2+2+1), device 1 gets 1,3,5,7 (2+2)
    assert FakeBatch.made == [(0, 2), (0, 2), (0, 1), (1, 2), (1, 2)]
    assert list(rows[0].keys()) == ["conf_num", "conf_rep", "cohesion", "seed", "effective_seed", "polarisation",
                                    "run_duration", "step_per_sec"]
    # no two runs share a stream, and (seed s, rep 1) is not (seed s + 1, rep 0)
    assert len({r["effective_seed"] for r in rows}) == len(rows)
    assert ex.run_seed(10, 0, 1) != ex.run_seed(11, 0, 0)


def test_configurations_with_their_own_world_form_their_own_batches(monkeypatch):
    """explore_parallel! builds an arbitrary state per configuration (model_exploration.rs:387-410):
    `dim` and `initial_flockers` may be swept; rows stay in run order"""
    import krabmaga_b200.explore as ex
    monkeypatch.setattr(ex, "FlockerBatch", FakeBatch)
    FakeBatch.made = []
    rows = ex.explore_parallel(5, 2, (100.0, 100.0), 50, 5.0,
                               {"cohesion": [1.0, 2.0, 3.0, 4.0], "initial_flockers": [50, 80, 50, 80],
                                "dim": [(100.0, 100.0), (100.0, 100.0), (200.0, 100.0), (100.0, 100.0)]},
                               mode=ExploreMode.Matched, devices=(0,), max_replicas_per_batch=8)
    assert [(r["conf_num"], r["conf_rep"]) for r in rows] == [(i, k) for i in range(4) for k in range(2)]
    assert [r["initial_flockers"] for r in rows[::2]] == [50, 80, 50, 80]
    # three distinct (dim, n) groups: conf 0 alone, confs 1 and 3 together, conf 2 alone
    assert sorted(FakeBatch.made) == [(0, 2), (0, 2), (0, 4)]
    for r in rows:
        assert r["polarisation"] == pytest.approx(expected_polarisation(r["cohesion"], r["effective_seed"]))


DIST_WORKER = r'''
import json, os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch.distributed as dist
import krabmaga_b200.explore as ex
from test_explore_host import FakeBatch
ex.FlockerBatch = FakeBatch
dist.init_process_group("gloo")
rank = dist.get_rank()
rows = ex.explore_distributed(4, 2, (100.0, 100.0), 30, 5.0, {"cohesion": [1.0, 2.0, 3.0, 4.0, 5.0]},
                              mode=ex.ExploreMode.Matched, device=rank)
line = {"rank": rank, "batches": FakeBatch.made,
        "rows": None if rows is None else [[r["conf_num"], r["conf_rep"], r["cohesion"], r["polarisation"]] for r in rows]}
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("RESULT " + json.dumps(line) + "\n")
'''


def test_explore_distributed_gloo_world_size_2(tmp_path):
    """explore_distributed_mpi! (explore/mpi/model_exploration.rs:121-290): configuration c runs on
    rank c % world with all its repetitions; the root gathers every row"""
    import json
    import os
    import socket
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(DIST_WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), root],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    res = {}
    for ln in out.stdout.splitlines():
        if "RESULT " in ln:
            d = json.loads(ln.split("RESULT ", 1)[1])
            res[d["rank"]] = d
    assert sorted(res) == [0, 1]
    assert res[1]["rows"] is None
    rows = res[0]["rows"]
    assert [(r[0], r[1]) for r in rows] == [(c, k) for c in range(5) for k in range(2)]
    assert [r[2] for r in rows[::2]] == [1.0, 2.0, 3.0, 4.0, 5.0]
    import krabmaga_b200.explore as ex
    for c, k, coh, pol in rows:
        assert pol == pytest.approx(expected_polarisation(coh, ex.run_seed(42, c, k)))
    # rank 0 ran configurations 0, 2, 4 (6 runs, one batch on its device), rank 1 ran 1, 3 (4 runs)
    assert res[0]["batches"] == [[0, 6]] and res[1]["batches"] == [[1, 4]]
