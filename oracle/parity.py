"""Test infrastructure (like everything under oracle/): which episodes of an Approach -> Finisher evaluation are "within tolerance
of a threshold" (BASELINE.json north_star) -- decided by the fp64 oracle itself, not by the code under test.

Only tests/, __graft_entry__.smoke() and bench.py's parity leg import this module; the product never does.
"""

from __future__ import annotations

import numpy as np

from . import kin_oracle as ko

def scale_decision_thresholds(cfg, factor: float):
    """The config with every threshold a success / handoff DECISION of the Approach -> Finisher evaluation compares against scaled by
    ``factor``: near-goal zone (dwell counter), dock-coarse-ready and finisher-ready predicates (handoff), termination success pose.
    Used to decide whether an episode is "within tolerance of a threshold": its fp64 oracle outcome changes under a small scaling."""
    from dataclasses import replace

    rc, tc = cfg.reward_config, cfg.termination_config
    names = ("near_goal_pos_threshold_m", "near_goal_ori_threshold_rad", "pre_near_goal_pos_threshold_m",
             "dock_coarse_ready_pos_threshold_m", "dock_coarse_ready_ori_threshold_rad", "dock_coarse_ready_action_threshold",
             "dock_coarse_ready_dq_threshold", "finisher_ready_pos_threshold_m", "finisher_ready_ori_threshold_rad",
             "finisher_ready_action_threshold", "finisher_ready_dq_threshold")
    rc2 = replace(rc, **{k: getattr(rc, k) * factor for k in names})
    tc2 = replace(tc, success_pos_threshold_m=tc.success_pos_threshold_m * factor, success_ori_threshold_rad=tc.success_ori_threshold_rad * factor)
    return replace(cfg, reward_config=rc2, termination_config=tc2)


def threshold_sensitive_episodes(approach_cfg, finisher_cfg, approach_policy, finisher_policy, suite, band: float, n_threads: int = 0, **kw):
    """(nominal oracle result, bool[n]): episodes whose fp64 oracle success flag is NOT invariant under scaling every decision threshold
    by (1 - band) and (1 + band) -- the operational meaning of north_star's "episodes within tolerance of a threshold"."""
    def run(f):
        a = approach_cfg if f == 1.0 else scale_decision_thresholds(approach_cfg, f)
        b = finisher_cfg if f == 1.0 else scale_decision_thresholds(finisher_cfg, f)
        res, _ = ko.eval_approach_finisher(ko.params_from_config(a), ko.params_from_config(b), approach_policy, finisher_policy,
                                           initial_q=np.asarray(suite.initial_q, np.float32).astype(float),
                                           goal_q=None if suite.goal_q is None else np.asarray(suite.goal_q, np.float32).astype(float),
                                           goal_pose6=None if getattr(suite, "goal_pose6", None) is None else np.asarray(suite.goal_pose6, np.float32).astype(float),
                                           initial_dq=None if getattr(suite, "initial_dq", None) is None else np.asarray(suite.initial_dq, np.float32).astype(float),
                                           initial_prev_action=None if getattr(suite, "initial_prev_action", None) is None else np.asarray(suite.initial_prev_action, np.float32).astype(float),
                                           n_threads=n_threads, **kw)
        return res
    nominal, lo, hi = run(1.0), run(1.0 - band), run(1.0 + band)
    sensitive = (lo["success"] != nominal["success"]) | (hi["success"] != nominal["success"])
    return nominal, sensitive
