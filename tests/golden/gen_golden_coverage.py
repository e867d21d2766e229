#!/usr/bin/env python
"""Golden fixture for the full-workspace coverage summary, produced by RUNNING THE LIVE REFERENCE (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden_coverage.py

Calls the reference's own helpers (eval/eval_full_workspace_coverage.py: _select_pairs, _run_pairs, _summarize, _bucket_metrics;
workspace/adaptive_frontier_sampler.py: update_bucket_priorities) with the bundled random-start + finisher checkpoints, seed 940001,
96 episodes per split -- the evaluation behind the published 0.802 / 0.240 / 0.219 table.  Output: tests/golden/coverage_summary.json
"""
from __future__ import annotations

import json
import sys
import tempfile
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import gen_golden as gg  # noqa: E402

from hrl_trainer.kinematic_phase1.workspace.adaptive_frontier_sampler import update_bucket_priorities  # noqa: E402

ref_cov = gg.ref_cov


def main() -> None:
    policies, _ = gg.load_checkpoints()
    cfgs = gg.merged_configs()
    cfg_path = gg.config_dir() / "workspace_full_coverage_randomstart_overnight.yaml"
    approach_cfg = gg.to_env_config(cfgs["randomstart_overnight"])
    finisher_cfg = gg.to_env_config(cfgs["finisher_noop_ft"])
    seed = 940001
    rng = np.random.default_rng(seed)
    targets, target_summary = ref_cov.generate_workspace_target_map(config_path=cfg_path, seed=seed + 1, stage_samples_per_stage=96, random_samples=384)
    starts, _ = ref_cov.generate_workspace_start_state_map(config_path=cfg_path, seed=seed + 2, stage_samples_per_stage=48, random_samples=384)
    with tempfile.TemporaryDirectory() as td:
        d = Path(td)
        ref_cov.write_target_map(targets, {}, d)
        ref_cov.write_start_state_map(starts, {}, d)
        pairs, _ = ref_cov.build_pair_sampler_summary(start_map_path=d / "start_state_map.jsonl", target_map_path=d / "target_map.jsonl",
                                                      seed=seed + 3, pair_count=2048)
        starts_by_id = {r["start_id"]: r for r in ref_cov.load_jsonl(d / "start_state_map.jsonl")}
        targets_by_id = {r["target_id"]: r for r in ref_cov.load_jsonl(d / "target_map.jsonl")}
    split_rows = {}
    for split in ("known", "frontier", "stress"):
        sel = ref_cov._select_pairs(pairs, mode=split, limit=96, rng=rng)
        split_rows[split] = ref_cov._run_pairs(pairs=sel, starts_by_id=starts_by_id, targets_by_id=targets_by_id, approach_model=policies["randomstart"],
                                               approach_env_cfg=approach_cfg, finisher_model=policies["finisher"], finisher_env_cfg=finisher_cfg,
                                               handoff_confirm_steps=2)
        print(split, ref_cov._summarize(split_rows[split])["success_rate"])
    all_rows = [r for rows in split_rows.values() for r in rows]
    metrics = ref_cov._bucket_metrics(all_rows)
    priorities = update_bucket_priorities(metrics)
    out = {
        "seed": seed, "target_bucket_ids": [t.bucket_id for t in targets], "target_bucket_count": target_summary["bucket_count"],
        "random_start_known_workspace": ref_cov._summarize(split_rows["known"]),
        "random_start_frontier": ref_cov._summarize(split_rows["frontier"]),
        "full_reachable_stress": ref_cov._summarize(split_rows["stress"]),
        "bucket_metrics": metrics, "priorities": [p.__dict__ for p in priorities],
        "row_failure_reasons": {k: [r["failure_reason"] for r in rows] for k, rows in split_rows.items()},
        "row_success": {k: [bool(r["success"]) for r in rows] for k, rows in split_rows.items()},
    }
    (gg.GOLD / "coverage_summary.json").write_text(json.dumps(out))
    print("wrote coverage_summary.json:", len(metrics), "buckets")


if __name__ == "__main__":
    main()
