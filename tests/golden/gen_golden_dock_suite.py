#!/usr/bin/env python
"""Golden dock (Finisher) eval suite, produced by RUNNING THE LIVE REFERENCE (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden_dock_suite.py

eval/fixed_eval_suite.py::build_dock_eval_suite on the finisher config (close-bucket probability 0.15, handoff buffer absent
upstream), seed 700001, 24 episodes.  Writes tests/golden/dock_suite.npz.
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import gen_golden as gg  # noqa: E402
from hrl_trainer.kinematic_phase1.eval.fixed_eval_suite import build_dock_eval_suite  # noqa: E402

cfg = gg.to_env_config(gg.merged_configs()["finisher_noop_ft"])
eps = build_dock_eval_suite(cfg, seed=700001, n_episodes=24)
np.savez_compressed(Path(__file__).resolve().parent / "dock_suite.npz", initial_q=np.array([e.initial_q for e in eps]),
                    goal_q=np.array([e.goal_q for e in eps]),
                    goal_pose6=np.array([e.goal_pose6 if e.goal_pose6 is not None else [np.nan] * 6 for e in eps], dtype=float))
print("dock_suite.npz", len(eps))
