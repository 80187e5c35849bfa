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
