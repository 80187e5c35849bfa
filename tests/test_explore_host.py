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
