#!/usr/bin/env python
"""Golden result of the reference's sequential route evaluation, produced by RUNNING THE LIVE REFERENCE (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden_route_eval.py

Chains eval/eval_route_curriculum.py::_roll_one over the synthetic 40-waypoint route of trace_route.npz with the bundled
route_prefix120 checkpoint (plain-torch stand-in for SB3 predict), then the reference's own _summarize_rows / _chunk_metrics /
_failure_reason.  Writes tests/golden/route_eval.json (per-waypoint rows, summary, chunk metrics, failure report).
"""
from __future__ import annotations

import json
import sys
import tempfile
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import gen_golden as gg  # noqa: E402  (sets up sys.path for the reference and the policy stand-in)
from hrl_trainer.kinematic_phase1.eval import eval_route_curriculum as erc  # noqa: E402


def main() -> None:
    policies, _ = gg.load_checkpoints()
    cfgs = gg.merged_configs()
    cfg = cfgs["route_prefix120"]
    route_q = gg.synthetic_route()
    with tempfile.TemporaryDirectory() as td:
        path = Path(td) / "route.json"
        path.write_text(json.dumps({"route_q": route_q.tolist()}))
        route = gg.load_route_dataset(path)
    env_cfg = gg.RouteEnvConfig(base_env_config=gg.to_env_config(cfg), reset_config=gg.RouteResetSamplerConfig(max_route_index=len(route) - 1),
                                reward_config=gg.route_reward_cfg(cfg), observation_config=gg.RouteObservationConfig(include_route_keys=True))
    env = gg.RouteKinematicEnv(route=route, config=env_cfg)
    model = policies["route_prefix120"]
    rows = []
    cq, cdq, cpa = route.waypoint(0).q_goal.copy(), np.zeros(7), np.zeros(7)
    for idx in range(1, len(route)):
        row = erc._roll_one(env, model, initial_q=cq, goal_index=idx, initial_dq=cdq, initial_prev_action=cpa)
        rows.append({k: v for k, v in row.items() if k not in {"final_q", "final_dq", "final_prev_action"}})
        cq, cdq, cpa = np.asarray(row["final_q"], float), np.asarray(row["final_dq"], float), np.asarray(row["final_prev_action"], float)
    summary = erc._summarize_rows(rows, route)
    failure = next((r for r in rows if not r["success"]), None)
    out = {"rows": rows, "summary": summary, "chunk_metrics": erc._chunk_metrics(rows),
           "failure_report": {"first_failure_index": summary["first_failure_index"], "first_failure_reason": summary["first_failure_reason"],
                              "first_failure": failure},
           "failure_reasons": [None if r["success"] else erc._failure_reason(r) for r in rows]}
    (Path(__file__).resolve().parent / "route_eval.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    print("route_eval.json:", len(rows), "rows, success_rate", summary["success_rate"], "prefix", summary["longest_success_prefix"],
          "first failure", summary["first_failure_index"], summary["first_failure_reason"])


if __name__ == "__main__":
    main()
