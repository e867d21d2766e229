"""CPU tests of the coverage tooling's host logic against outputs of the live reference (tests/golden/coverage_summary.json)."""

from __future__ import annotations

import json

from rl_brain_trainer_b200 import coverage, workspace

from ._util import GOLD, env_config


def _gold():
    return json.loads((GOLD / "coverage_summary.json").read_text())


def test_target_bucket_ids_match_reference():
    g = _gold()
    cfg = env_config("randomstart_overnight")
    targets = workspace.generate_workspace_target_map(cfg, seed=940001 + 1, stage_samples_per_stage=96, random_samples=384)
    b = coverage.target_buckets(targets)
    assert [b.ids[c] for c in b.code] == g["target_bucket_ids"]
    assert b.count == g["target_bucket_count"] == len(set(g["target_bucket_ids"]))


def test_bucket_priorities_match_reference():
    g = _gold()
    got = coverage.update_bucket_priorities(g["bucket_metrics"])
    assert [p.bucket_id for p in got] == [p["bucket_id"] for p in g["priorities"]]
    for a, b in zip(got, g["priorities"]):
        assert a.category == b["category"] and abs(a.sampling_priority - b["sampling_priority"]) < 1e-12 and a.failure_count == b["failure_count"]
    # the categories' decision table (adaptive_frontier_sampler.py:21-39)
    c = coverage.classify_bucket
    assert c(success_rate=0.5, mean_min_error=0.01, mean_final_error=0.01, previous_success_rate=0.8) == "forgetting_risk"
    assert c(success_rate=0.9, mean_min_error=0.0, mean_final_error=0.0) == "mastered"
    assert c(success_rate=0.35, mean_min_error=0.0, mean_final_error=0.0) == "frontier"
    assert c(success_rate=0.1, mean_min_error=0.03, mean_final_error=0.05) == "too_hard"
    assert c(success_rate=0.3, mean_min_error=0.01, mean_final_error=0.02) == "hard_but_promising"
    assert c(success_rate=0.3, mean_min_error=0.02, mean_final_error=0.02) == "stress"
    assert coverage.priority_for_category("frontier") == 1.0 and coverage.priority_for_category("unknown") == 0.20
