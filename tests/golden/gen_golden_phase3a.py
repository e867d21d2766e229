#!/usr/bin/env python
"""Golden vectors for the Phase-3A Gazebo bridge helpers, produced by RUNNING THE LIVE REFERENCE (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden_phase3a.py

Drives the module-level functions hrl_trainer/v5/phase3a_controlled_sim.py imports (`:22-32`) -- compute_ee_pose6,
build_observation, pose_error_components, clip_joint_configuration, delta_limits -- and the bridge's own
action_to_command_q / effective_action_delta_scale (`:131-160`) on seeded inputs.  Writes tests/golden/phase3a.npz.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/hrl_ws/src/hrl_trainer")
sys.path.insert(0, str(REF))
from hrl_trainer.kinematic_phase1.envs.observation_builder import build_observation  # noqa: E402
from hrl_trainer.kinematic_phase1.kinematics.fk_interface import compute_ee_pose6  # noqa: E402
from hrl_trainer.kinematic_phase1.kinematics.joint_limits import clip_joint_configuration, default_joint_specs, delta_limits  # noqa: E402
from hrl_trainer.kinematic_phase1.kinematics.pose_utils import pose_error_components  # noqa: E402

KEYS = sorted(["q", "dq", "prev_action", "goal_pos_err", "goal_ori_err", "wp_pos_err", "wp_ori_err", "next_wp_pos_err", "next_wp_ori_err",
               "task_type", "mode_flag", "progress", "joint_limit_margin"])


def main() -> None:
    rng = np.random.default_rng(20260117)
    specs = default_joint_specs()
    lo = np.array([s.lower for s in specs])
    hi = np.array([s.upper for s in specs])
    n = 48
    q = rng.uniform(lo - 0.05, hi + 0.05, (n, 7))                 # some outside the limits
    q_goal = rng.uniform(lo, hi, (n, 7))
    dq = rng.uniform(-0.06, 0.06, (n, 7))
    prev_action = rng.uniform(-1.3, 1.3, (n, 7))
    ep = rng.uniform(-0.2, 1.2, n)
    dw = rng.uniform(-0.2, 1.4, n)
    mode = rng.integers(0, 4, n)
    wp_q = rng.uniform(lo, hi, (n, 7))
    nwp_q = rng.uniform(lo, hi, (n, 7))
    use_wp = rng.random(n) < 0.5
    cur = np.stack([compute_ee_pose6(x) for x in q])
    goal = np.stack([compute_ee_pose6(x) for x in q_goal])
    wp = np.stack([compute_ee_pose6(x) for x in wp_q])
    nwp = np.stack([compute_ee_pose6(x) for x in nwp_q])
    obs = np.zeros((n, 56), dtype=np.float32)
    perr = np.zeros((n, 3))
    oerr = np.zeros((n, 3))
    for i in range(n):
        o = build_observation(q=q[i], dq=dq[i], prev_action=prev_action[i], current_pose6=cur[i], goal_pose6=goal[i], joint_specs=specs,
                              episode_progress=float(ep[i]), dwell_progress=float(dw[i]), mode_index=int(mode[i]),
                              current_waypoint_pose6=wp[i] if use_wp[i] else None, next_waypoint_pose6=nwp[i] if use_wp[i] else None)
        obs[i] = np.concatenate([np.asarray(o[k], dtype=np.float32).reshape(-1) for k in KEYS])
        perr[i], oerr[i] = pose_error_components(cur[i], goal[i])
    clipped = np.stack([clip_joint_configuration(x, specs) for x in q])
    np.savez_compressed(Path(__file__).resolve().parent / "phase3a.npz", q=q, q_goal=q_goal, dq=dq, prev_action=prev_action, episode_progress=ep,
                        dwell_progress=dw, mode_index=mode, use_wp=use_wp, current_pose6=cur, goal_pose6=goal, wp_pose6=wp, next_wp_pose6=nwp,
                        obs56=obs, pos_err=perr, ori_err=oerr, clipped_q=clipped, delta_limits=np.asarray(delta_limits(specs), dtype=float))
    print("wrote phase3a.npz", obs.shape)


if __name__ == "__main__":
    main()
