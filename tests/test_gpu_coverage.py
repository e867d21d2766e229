"""evaluate_full_workspace_coverage on the GPU against the live reference's summary of the same evaluation (seed 940001, 96 per split)."""

from __future__ import annotations

import json

import numpy as np
import pytest

from ._util import GOLD, env_config

pytestmark = pytest.mark.gpu


def test_full_workspace_coverage_matches_reference_summary():
    from rl_brain_trainer_b200 import coverage
    from rl_brain_trainer_b200.policy import PolicyWeights

    g = json.loads((GOLD / "coverage_summary.json").read_text())
    acfg, fcfg = env_config("randomstart_overnight"), env_config("finisher_noop_ft")
    out = coverage.evaluate_full_workspace_coverage(acfg, PolicyWeights.preset("randomstart", "cuda"), fcfg, PolicyWeights.preset("finisher", "cuda"))
    for split in ("random_start_known_workspace", "random_start_frontier", "full_reachable_stress"):
        a, b = out[split], g[split]
        assert a["episode_count"] == b["episode_count"] == 96
        # fp32 vs fp64 may flip an episode that ends within rounding of a threshold: +-3 of 96
        assert abs(a["success_rate"] - b["success_rate"]) <= 3 / 96 + 1e-9, split
        assert abs(a["ready_rate"] - b["ready_rate"]) <= 3 / 96 + 1e-9 and abs(a["dwell_success_rate"] - b["dwell_success_rate"]) <= 3 / 96 + 1e-9
        # the pairs themselves are identical (host PCG64 port): distances and per-source counts are exact
        assert abs(a["average_start_target_joint_distance"] - b["average_start_target_joint_distance"]) < 1e-9
        assert abs(a["average_start_target_ee_distance"] - b["average_start_target_ee_distance"]) < 1e-6
        counts = lambda d: {k: v["episode_count"] for k, v in d["success_by_start_source"].items()}  # noqa: E731
        assert counts(a) == counts(b)
        assert abs(a["mean_final_position_error"] - b["mean_final_position_error"]) < 0.15 * b["mean_final_position_error"] + 1e-4
        ra, rb = a["failure_reason_counts"], b["failure_reason_counts"]
        assert sum(ra.values()) == 96 and set(ra) <= set(rb) | {"success", "position", "orientation", "motion_action", "motion_dq", "dwell",
                                                                 "timeout_or_regression"}
        assert sum(abs(ra.get(k, 0) - rb.get(k, 0)) for k in set(ra) | set(rb)) <= 8, (split, ra, rb)
    # buckets: same episodes in the same buckets; rates within the flip band
    assert out["total_eval_bucket_count"] == len(g["bucket_metrics"])
    for bid, m in g["bucket_metrics"].items():
        mine = out["bucket_metrics"][bid]
        assert mine["episode_count"] == m["episode_count"]
        assert abs(mine["failure_count"] - m["failure_count"]) <= 3
        assert abs(mine["mean_min_position_error"] - m["mean_min_position_error"]) < 1e-3
    stable = sum(1 for m in g["bucket_metrics"].values() if m["success_rate"] >= 0.85)
    assert abs(out["stable_bucket_fraction"] - stable / len(g["bucket_metrics"])) <= 3 / len(g["bucket_metrics"])
    top_ref = {p["bucket_id"] for p in g["priorities"][:10]}
    top = {p["bucket_id"] for p in out["top_sampling_priorities"][:10]}
    assert len(top & top_ref) >= 7
    assert out["env_steps"] > 3 * 96 * 100


def test_bucket_metrics_at_scale_are_consistent():
    """10^5 random-start pairs: the device segment reductions equal a numpy group-by of the same rows; previous rates are carried."""
    import torch

    from rl_brain_trainer_b200 import coverage, workspace as ws
    from rl_brain_trainer_b200.policy import PolicyWeights
    from rl_brain_trainer_b200.rollout import VARIANT_TC, ApproachFinisherRollout

    acfg, fcfg = env_config("randomstart_overnight"), env_config("finisher_noop_ft")
    targets = ws.generate_workspace_target_map(acfg, seed=11, stage_samples_per_stage=32, random_samples=128)
    starts = ws.generate_workspace_start_state_map(acfg, seed=12, stage_samples_per_stage=16, random_samples=128)
    pairs = ws.build_pair_table(starts, targets, seed=13, pair_count=100_000)
    buckets = coverage.target_buckets(targets)
    ro = ApproachFinisherRollout(acfg, PolicyWeights.preset("randomstart", "cuda"), fcfg, PolicyWeights.preset("finisher", "cuda"), variant=VARIANT_TC)
    res = ro.evaluate_suite(ws.pairs_to_suite(starts, targets, pairs, np.arange(100_000)))
    code = buckets.code[pairs.target]
    m = coverage.bucket_metrics(res, torch.as_tensor(code, device="cuda"), buckets)
    ok, fin = res.success.cpu().numpy(), res.final_position_error.cpu().numpy().astype(np.float64)
    assert sum(v["episode_count"] for v in m.values()) == 100_000 and len(m) == len(np.unique(code))
    for b in np.unique(code)[:40]:
        sel = code == b
        v = m[buckets.ids[b]]
        assert v["episode_count"] == int(sel.sum()) and v["failure_count"] == int((~ok[sel]).sum())
        assert abs(v["success_rate"] - ok[sel].mean()) < 1e-12 and abs(v["mean_final_position_error"] - fin[sel].mean()) < 1e-9
    m2 = coverage.bucket_metrics(res, torch.as_tensor(code, device="cuda"), buckets, previous=m)
    assert all(abs(v["previous_success_rate"] - m[k]["success_rate"]) < 1e-15 for k, v in m2.items())
    pr = coverage.update_bucket_priorities(m2)
    assert len(pr) == len(m2) and all(pr[i].sampling_priority >= pr[i + 1].sampling_priority for i in range(len(pr) - 1))
