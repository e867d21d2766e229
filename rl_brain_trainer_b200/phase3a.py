"""The helpers the reference's Phase-3A Gazebo bridge imports, served by the CUDA library.

``hrl_trainer/v5/phase3a_controlled_sim.py:22-32`` drives Gazebo with the frozen Approach / Finisher checkpoints and, per control
tick, calls ``compute_ee_pose6`` (fk_interface.py:21), ``build_observation`` (observation_builder.py:29-94),
``pose_error_components`` (pose_utils.py:19-26), ``clip_joint_configuration`` / ``delta_limits`` (joint_limits.py) and its own
``action_to_command_q`` / ``effective_action_delta_scale`` (:131-160).  A maintainer points those imports at this module and the
bridge runs against the same FK / observation arithmetic the batched kinematic env uses (``kin_fk_pose6``, ``kin_env_observe``),
which is what makes the sim-to-Gazebo parity check meaningful: the observation a checkpoint sees in Gazebo is produced by the
code that produced its training observations.  ROS / Gazebo code itself is untouched (north_star).

Everything here is a 1-env, host-synchronous call (the bridge ticks at a few Hz): conformance, not throughput.  The joint-limit
helpers are plain config arithmetic on 7 numbers and stay on the host like the reference's.
"""

from __future__ import annotations

from typing import Any, Mapping, Sequence

import numpy as np
import torch

from . import _lib
from .config import JOINT_ORDER, JointSpec, Phase1EnvConfig, default_joint_specs, load_yaml_file, to_env_config
from .env import OBS_KEYS, OBS_SLICES, BatchedArmKinematicEnv

_D = _lib.define

__all__ = ["JOINT_ORDER", "JointSpec", "Phase3AKinematics", "action_to_command_q", "build_observation", "clip_joint_configuration",
           "compute_ee_pose6", "default_joint_specs", "delta_limits", "effective_action_delta_scale", "load_yaml_file", "pose_error_components",
           "to_env_config"]


def clip_joint_configuration(q: Sequence[float], joint_specs: Sequence[JointSpec]) -> np.ndarray:
    """joint_limits.py: element-wise clip to [lower, upper]."""
    q_arr = np.asarray(q, dtype=float)
    return np.clip(q_arr, np.array([s.lower for s in joint_specs], dtype=float), np.array([s.upper for s in joint_specs], dtype=float))


def delta_limits(joint_specs: Sequence[JointSpec]) -> np.ndarray:
    return np.array([s.delta_limit for s in joint_specs], dtype=float)


class Phase3AKinematics:
    """One 1-env device context: FK and the observation builder of the kinematic env for externally supplied joint states."""

    def __init__(self, config: Phase1EnvConfig | None = None, device: str | torch.device = "cuda") -> None:
        self.config = config or Phase1EnvConfig()
        self._b = BatchedArmKinematicEnv(self.config, 1, device, host_sampler=True)
        self._b.reset(options={"initial_q": np.zeros(7), "goal_q": np.zeros(7)})

    def compute_ee_pose6(self, q: Sequence[float]) -> np.ndarray:
        """``compute_ee_pose6`` (fk_interface.py:21 -> v5_1/ee_fk.py:120) through ``kin_fk_pose6``."""
        q_arr = np.asarray(q, dtype=float)
        if q_arr.shape != (7,):
            raise ValueError(f"Expected q shape (7,), got {q_arr.shape}")
        return self._b.fk_pose6_host(q_arr[None])[0]

    def _set_rows(self, row: int, values: Sequence[float]) -> None:
        v = torch.as_tensor(np.asarray(values, dtype=np.float32), device=self._b.device)
        self._b.state[row:row + v.numel(), 0] = v

    def _observe(self, goal_pose6: np.ndarray) -> np.ndarray:
        self._set_rows(_D("KIN_ROW_GOAL_POSE"), goal_pose6)
        return self._b.current_observation()[0].cpu().numpy()

    def pose_error_components(self, current_pose6: Sequence[float], goal_pose6: Sequence[float]) -> tuple[np.ndarray, np.ndarray]:
        """pose_utils.py:19-26: (goal - current) position, per-component wrapped Euler difference."""
        cur, goal = np.asarray(current_pose6, dtype=float), np.asarray(goal_pose6, dtype=float)
        d = goal[3:] - cur[3:]
        return goal[:3] - cur[:3], (d + np.pi) % (2.0 * np.pi) - np.pi

    def build_observation(self, *, q: Sequence[float], dq: Sequence[float], prev_action: Sequence[float], current_pose6: Sequence[float],
                          goal_pose6: Sequence[float], joint_specs: Sequence[JointSpec] | None = None, episode_progress: float,
                          dwell_progress: float, mode_index: int, current_waypoint_pose6: Sequence[float] | None = None,
                          next_waypoint_pose6: Sequence[float] | None = None, config: Any = None) -> dict[str, np.ndarray]:
        """``build_observation`` (observation_builder.py:29-94) for one externally measured joint state, through ``kin_env_observe``.

        ``joint_specs`` / ``config`` must be the ones this context was built with (the device parameter block already holds
        them); they are accepted for signature compatibility.
        """
        if joint_specs is not None and tuple(joint_specs) != tuple(self.config.joint_specs):
            raise ValueError("joint_specs differ from the ones this Phase3AKinematics was constructed with")
        if config is not None and (float(config.pos_err_scale_m) != float(self.config.observation_config.pos_err_scale_m) or
                                   float(config.ori_err_scale_rad) != float(self.config.observation_config.ori_err_scale_rad)):
            raise ValueError("observation config differs from the one this Phase3AKinematics was constructed with")
        self._set_rows(_D("KIN_ROW_Q"), q)
        self._set_rows(_D("KIN_ROW_DQ"), dq)
        self._set_rows(_D("KIN_ROW_PREV_ACTION"), prev_action)
        self._set_rows(_D("KIN_ROW_EE_POSE"), current_pose6)
        flags = int(np.clip(int(mode_index), 0, 3)) << _D("KIN_FLAG_MODE_SHIFT")
        self._b.state[_D("KIN_ROW_FLAGS"), 0] = torch.tensor(flags, dtype=torch.int32).view(torch.float32)
        flat = self._observe(np.asarray(goal_pose6, dtype=float))
        obs = {k: flat[OBS_SLICES[k]].copy() for k in OBS_KEYS}
        # waypoint errors are the goal-error arithmetic against another target pose (always absent on the Approach/Finisher path)
        for key, pose in (("wp", current_waypoint_pose6), ("next_wp", next_waypoint_pose6)):
            if pose is not None:
                other = self._observe(np.asarray(pose, dtype=float))
                obs[f"{key}_pos_err"] = other[OBS_SLICES["goal_pos_err"]].copy()
                obs[f"{key}_ori_err"] = other[OBS_SLICES["goal_ori_err"]].copy()
        # the bridge supplies its own progress fractions (tick / (episode_steps - 1), dwell / 5): the env's counters do not apply
        obs["progress"] = np.array([float(np.clip(episode_progress, 0.0, 1.0)), float(np.clip(dwell_progress, 0.0, 1.0)), 0.0], dtype=np.float32)
        return obs


_default: Phase3AKinematics | None = None


def _ctx() -> Phase3AKinematics:
    global _default
    if _default is None:
        _default = Phase3AKinematics()
    return _default


def compute_ee_pose6(q: Sequence[float]) -> np.ndarray:
    return _ctx().compute_ee_pose6(q)


def pose_error_components(current_pose6: Sequence[float], goal_pose6: Sequence[float]) -> tuple[np.ndarray, np.ndarray]:
    return _ctx().pose_error_components(current_pose6, goal_pose6)


def build_observation(**kwargs: Any) -> dict[str, np.ndarray]:
    return _ctx().build_observation(**kwargs)


def action_to_command_q(*, q: np.ndarray, action: np.ndarray, joint_specs: Sequence[JointSpec], action_delta_scale: float) -> np.ndarray:
    """phase3a_controlled_sim.py:131-140 -- the joint command the env's step would execute for this action."""
    a = np.clip(np.asarray(action, dtype=float).reshape(-1), -1.0, 1.0)
    return clip_joint_configuration(np.asarray(q, dtype=float) + a * delta_limits(joint_specs) * float(action_delta_scale), joint_specs)


def effective_action_delta_scale(phase_cfg: Any, pos_error: float) -> float:
    """phase3a_controlled_sim.py:143-160 (= ``_dynamic_action_delta_scale`` of the env, arm_kinematic_env.py:489-511)."""
    get = (lambda k: phase_cfg[k]) if isinstance(phase_cfg, Mapping) else (lambda k: getattr(phase_cfg, k))
    base = float(get("action_delta_scale"))
    if not get("dynamic_action_delta_scale_enabled"):
        return base
    near, far = float(get("dynamic_action_delta_scale_near_pos_threshold_m")), float(get("dynamic_action_delta_scale_far_pos_threshold_m"))
    if near <= 0.0 or far <= near:
        return base
    nm, fm = float(get("dynamic_action_delta_scale_near_multiplier")), float(get("dynamic_action_delta_scale_far_multiplier"))
    if pos_error <= near:
        mult = nm
    elif pos_error >= far:
        mult = fm
    else:
        mult = nm + (pos_error - near) / max(far - near, 1e-9) * (fm - nm)
    return float(base * max(mult, 0.0))
