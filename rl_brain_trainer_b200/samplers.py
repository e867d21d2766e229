"""Host-side reset samplers and eval-suite builders that consume numpy's PCG64 stream exactly like the reference.

Seed parity matters for evaluation: the reference builds its suites and random starts from
``np.random.default_rng(seed)`` in a data-dependent call order (conditional draws, rejection loops).
These functions make the same generator calls in the same order, so a given seed yields bit-identical
start/goal joint vectors (checked against ``tests/golden/samplers.npz``).  They are used by the 1-env
adapter and to build eval suites that are then uploaded; the large-batch training path uses the
device-side Philox samplers in ``csrc/kin_state.cuh`` instead (validated distributionally).

Reference: ``kinematic_phase1/envs/curriculum.py:90-101``, ``envs/reset_samplers.py:168-515``,
``kinematics/joint_limits.py:124-137``, ``eval/fixed_eval_suite.py:39-105``.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Sequence

import numpy as np

from .config import JointSpec, Phase1EnvConfig

FkFn = Callable[[np.ndarray], np.ndarray]


@dataclass
class ResetSample:
    initial_q: np.ndarray
    goal_q: np.ndarray
    goal_pose6: np.ndarray | None = None
    initial_dq: np.ndarray | None = None
    initial_prev_action: np.ndarray | None = None
    stage_index: int | None = None


def _bounds(specs: Sequence[JointSpec]) -> tuple[np.ndarray, np.ndarray]:
    return np.array([s.lower for s in specs], dtype=float), np.array([s.upper for s in specs], dtype=float)


def clip_q(q: np.ndarray, specs: Sequence[JointSpec]) -> np.ndarray:
    lo, hi = _bounds(specs)
    return np.clip(np.asarray(q, dtype=float), lo, hi)


def sample_joint_configuration(rng: np.random.Generator, specs: Sequence[JointSpec], margin_fraction: float = 0.1) -> np.ndarray:
    lo, hi = _bounds(specs)
    margin = np.maximum((hi - lo) * margin_fraction, 1e-6)
    return rng.uniform(low=lo + margin, high=hi - margin, size=(len(specs),)).astype(float)


def sample_stage_joint_target(rng: np.random.Generator, base_q: Sequence[float], noise_q: Sequence[float], specs: Sequence[JointSpec]) -> np.ndarray:
    base = np.asarray(base_q, dtype=float)
    noise = np.asarray(noise_q, dtype=float)
    if np.any(noise > 0.0):  # the draw only happens for a noisy shell -- part of the stream contract
        base = base + rng.uniform(low=-noise, high=noise)
    return clip_q(base, specs)


def sample_workspace_stage_index(rng: np.random.Generator, current_stage_index: int, stage_count: int, config: dict[str, Any] | None) -> int:
    current = int(np.clip(current_stage_index, 0, max(stage_count - 1, 0)))
    cfg = dict(config or {})
    if not bool(cfg.get("enabled", False)) or current <= 0:
        return current
    ratios = [max(float(cfg.get(k, d)), 0.0) for k, d in (("current_stage_ratio", 0.50), ("previous_stage_ratio", 0.25),
                                                             ("old_workspace_replay_ratio", 0.20), ("failure_replay_ratio", 0.05))]
    total = sum(ratios)
    if total <= 0.0:
        return current
    draw = float(rng.random() * total)
    if draw < ratios[0]:
        return current
    draw -= ratios[0]
    if draw < ratios[1] and current > 0:
        low = max(int(cfg.get("previous_stage_min_index", 0)), 0)
        return int(rng.integers(low, max(current - 1, low) + 1))
    draw -= ratios[1]
    old_max = int(np.clip(int(cfg.get("old_workspace_max_stage_index", min(5, current))), 0, min(stage_count - 1, current)))
    if draw < ratios[2] and old_max >= 0:
        return int(rng.integers(0, old_max + 1))
    replay_max = max(min(old_max, current - 1), 0)
    return int(rng.integers(0, replay_max + 1)) if replay_max > 0 else current


_SOURCES = ("home", "old_success", "random_valid", "frontier", "failure_recovery", "stress")
_SOURCE_KEYS = ("home_start_ratio", "old_successful_start_ratio", "random_valid_q_start_ratio", "frontier_pair_ratio",
                "failure_recovery_start_ratio", "stress_start_ratio")
_SOURCE_DEFAULTS = (0.15, 0.25, 0.25, 0.20, 0.10, 0.05)


def _target_stage(rng: np.random.Generator, source: str, current: int, n: int, cfg: dict[str, Any]) -> int:
    if source in ("home", "old_success"):
        return int(rng.integers(0, int(np.clip(cfg.get("known_target_max_stage_index", min(7, current)), 0, n - 1)) + 1))
    if source == "frontier":
        lo = int(np.clip(cfg.get("frontier_target_min_stage_index", min(8, current)), 0, n - 1))
        return int(rng.integers(lo, int(np.clip(cfg.get("frontier_target_max_stage_index", current), lo, n - 1)) + 1))
    if source == "stress":
        lo = int(np.clip(cfg.get("stress_target_min_stage_index", min(8, current)), 0, n - 1))
        return int(rng.integers(lo, int(np.clip(cfg.get("stress_target_max_stage_index", n - 1), lo, n - 1)) + 1))
    return int(rng.integers(0, int(np.clip(cfg.get("mixed_target_max_stage_index", current), 0, n - 1)) + 1))


def sample_random_start_workspace_pair(rng: np.random.Generator, config: Phase1EnvConfig, stage_index: int, cfg: dict[str, Any]) -> ResetSample:
    specs = config.joint_specs
    stages = config.curriculum_config.stages
    n = len(stages)
    current = int(np.clip(stage_index, 0, n - 1))
    ratios = [max(float(cfg.get(k, d)), 0.0) for k, d in zip(_SOURCE_KEYS, _SOURCE_DEFAULTS)]
    total = sum(ratios)
    source = "old_success"
    if total > 0.0:
        draw = float(rng.random() * total)
        for name, value in zip(_SOURCES, ratios):
            if draw <= value:
                source = name
                break
            draw -= value
    tstage = _target_stage(rng, source, current, n, cfg)
    target_q = sample_stage_joint_target(rng, stages[tstage].goal_q, stages[tstage].goal_noise, specs)
    if source == "home":
        st = stages[min(int(cfg.get("home_stage_index", 0)), n - 1)]
        start_q = sample_stage_joint_target(rng, st.start_q, st.start_noise, specs)
    elif source == "old_success":
        st = stages[int(rng.integers(0, int(np.clip(cfg.get("old_success_max_stage_index", min(7, current)), 0, n - 1)) + 1))]
        start_q = sample_stage_joint_target(rng, st.goal_q, st.goal_noise, specs)
    elif source == "frontier":
        lo = int(np.clip(cfg.get("frontier_min_stage_index", min(8, current)), 0, n - 1))
        st = stages[int(rng.integers(lo, int(np.clip(cfg.get("frontier_max_stage_index", current), lo, n - 1)) + 1))]
        start_q = sample_stage_joint_target(rng, st.start_q, st.start_noise, specs)
    elif source == "failure_recovery":
        noise = np.asarray(cfg.get("failure_recovery_q_noise", [0.04] * len(specs)), dtype=float)
        start_q = clip_q(target_q + rng.uniform(-noise, noise), specs)
    elif source == "stress":
        start_q = sample_joint_configuration(rng, specs, float(cfg.get("stress_start_margin_fraction", config.start_sample_margin_fraction)))
    else:
        start_q = sample_joint_configuration(rng, specs, float(cfg.get("random_valid_start_margin_fraction", config.start_sample_margin_fraction)))
    dq_noise = np.asarray(cfg.get("initial_dq_noise", [0.0] * len(specs)), dtype=float)
    pa_noise = np.asarray(cfg.get("initial_prev_action_noise", [0.0] * len(specs)), dtype=float)
    initial_dq = rng.uniform(-dq_noise, dq_noise) if np.any(dq_noise > 0.0) else np.zeros(len(specs))
    initial_pa = rng.uniform(-pa_noise, pa_noise) if np.any(pa_noise > 0.0) else np.zeros(len(specs))
    min_l2 = float(cfg.get("min_pair_joint_l2", 0.0))
    if min_l2 > 0.0:
        for _ in range(12):
            if float(np.linalg.norm(target_q - start_q)) >= min_l2:
                break
            tstage = _target_stage(rng, source, current, n, cfg)
            target_q = sample_stage_joint_target(rng, stages[tstage].goal_q, stages[tstage].goal_noise, specs)
    return ResetSample(initial_q=clip_q(start_q, specs), goal_q=clip_q(target_q, specs), initial_dq=initial_dq,
                       initial_prev_action=initial_pa, stage_index=tstage)


def sample_approach_reset(rng: np.random.Generator, config: Phase1EnvConfig, stage_index: int) -> ResetSample:
    cur = config.curriculum_config
    specs = config.joint_specs
    rs_cfg = dict((config.workspace_stage_sampling or {}).get("random_start_pair_sampling", {}))
    if bool(rs_cfg.get("enabled", False)) and cur.enabled and cur.stages:
        return sample_random_start_workspace_pair(rng, config, stage_index, rs_cfg)
    if cur.enabled and cur.stages:
        idx = sample_workspace_stage_index(rng, stage_index, len(cur.stages), config.workspace_stage_sampling)
        st = cur.stages[idx]
        initial_q = sample_stage_joint_target(rng, st.start_q, st.start_noise, specs)
        goal_q = sample_stage_joint_target(rng, st.goal_q, st.goal_noise, specs)
        return ResetSample(initial_q=initial_q, goal_q=goal_q, stage_index=idx)
    initial_q = sample_joint_configuration(rng, specs, config.start_sample_margin_fraction)
    goal_q = sample_joint_configuration(rng, specs, config.goal_sample_margin_fraction)
    return ResetSample(initial_q=initial_q, goal_q=goal_q)


def _wrap(v: np.ndarray) -> np.ndarray:
    return (v + np.pi) % (2.0 * np.pi) - np.pi


def _close_bucket_initial_q(rng: np.random.Generator, config: Phase1EnvConfig, goal_q: np.ndarray, goal_pose6: np.ndarray, fk: FkFn) -> np.ndarray:
    """Near-success-but-not-yet states by rejection (reset_samplers.py:474-515), same draw order."""
    d = config.dock_reset_config
    noise = np.asarray(d.close_init_q_noise, dtype=float)
    best_q, best_dist = None, float("inf")
    for _ in range(max(int(d.close_bucket_max_attempts), 1)):
        cand = clip_q(goal_q + rng.uniform(low=-noise, high=noise), config.joint_specs)
        pose = np.asarray(fk(cand), dtype=float)
        pos = float(np.linalg.norm(goal_pose6[:3] - pose[:3]))
        ori = float(np.linalg.norm(_wrap(goal_pose6[3:] - pose[3:])))
        if d.close_bucket_min_pos_error_m <= pos <= d.close_bucket_max_pos_error_m and d.close_bucket_min_ori_error_rad <= ori <= d.close_bucket_max_ori_error_rad:
            return cand
        if pos < d.close_bucket_min_pos_error_m:
            dist = d.close_bucket_min_pos_error_m - pos
        elif pos > d.close_bucket_max_pos_error_m:
            dist = pos - d.close_bucket_max_pos_error_m
        else:
            dist = max(d.close_bucket_min_ori_error_rad - ori, ori - d.close_bucket_max_ori_error_rad, 0.0)
        if dist < best_dist:
            best_q, best_dist = cand, float(dist)
    return best_q if best_q is not None else clip_q(goal_q, config.joint_specs)


def sample_dock_reset(rng: np.random.Generator, config: Phase1EnvConfig, stage_index: int, fk: FkFn | None = None,
                      handoff_states: Sequence[dict[str, Any]] = ()) -> ResetSample:
    d = config.dock_reset_config
    cur = config.curriculum_config
    specs = config.joint_specs
    if d.handoff_state_probability > 0.0 and handoff_states and rng.random() < d.handoff_state_probability:
        st = handoff_states[int(rng.integers(len(handoff_states)))]
        return ResetSample(initial_q=np.asarray(st["initial_q"], dtype=float), goal_q=np.asarray(st["goal_q"], dtype=float),
                           goal_pose6=np.asarray(st["goal_pose6"], dtype=float),
                           initial_dq=np.asarray(st.get("initial_dq", [0.0] * 7), dtype=float),
                           initial_prev_action=np.asarray(st.get("initial_prev_action", [0.0] * 7), dtype=float))
    if cur.enabled and cur.stages:
        st = cur.stages[int(np.clip(stage_index, 0, len(cur.stages) - 1))]
        goal_q = sample_stage_joint_target(rng, st.goal_q, st.goal_noise, specs)
    else:
        goal_q = sample_stage_joint_target(rng, d.goal_q, d.goal_noise, specs)
    if d.close_bucket_probability > 0.0 and rng.random() < d.close_bucket_probability:
        if fk is None:
            raise ValueError("the close-bucket dock reset needs an FK callback")
        goal_pose6 = np.asarray(fk(goal_q), dtype=float)
        return ResetSample(initial_q=_close_bucket_initial_q(rng, config, goal_q, goal_pose6, fk), goal_q=goal_q)
    noise = np.asarray(d.init_q_noise, dtype=float)
    return ResetSample(initial_q=clip_q(goal_q + rng.uniform(low=-noise, high=noise), specs), goal_q=goal_q)


def sample_reset(rng: np.random.Generator, config: Phase1EnvConfig, mode_name: str, stage_index: int, fk: FkFn | None = None,
                 handoff_states: Sequence[dict[str, Any]] = ()) -> ResetSample:
    """Dispatch of ``ArmKinematicEnv.reset`` without explicit ``initial_q`` (arm_kinematic_env.py:157-176)."""
    if mode_name in ("dock", "dock_coarse"):
        return sample_dock_reset(rng, config, stage_index, fk, handoff_states)
    return sample_approach_reset(rng, config, stage_index)


# ------------------------------------------------------------------------------------------------
# eval suites (eval/fixed_eval_suite.py)
# ------------------------------------------------------------------------------------------------
@dataclass
class EvalSuite:
    """Column-major eval suite: one row per episode (``EvalEpisodeSpec`` of the reference, stacked)."""

    initial_q: np.ndarray           # [n,7]
    goal_q: np.ndarray              # [n,7]
    goal_pose6: np.ndarray | None = None      # [n,6]; None -> the device computes FK(goal_q)
    initial_dq: np.ndarray | None = None
    initial_prev_action: np.ndarray | None = None

    def __len__(self) -> int:
        return int(self.initial_q.shape[0])


def build_curriculum_local_eval_suite(config: Phase1EnvConfig, *, seed: int = 700001, stage_index: int = 0, n_episodes: int = 10) -> EvalSuite:
    """Same stream as ``fixed_eval_suite.py:78-105``: per episode a start-noise draw (if any) then a goal-noise draw."""
    cur = config.curriculum_config
    if not cur.enabled:
        raise ValueError("Curriculum-local eval requires curriculum to be enabled")
    if not cur.stages:
        raise ValueError("No curriculum stages are defined")
    st = cur.stages[int(np.clip(stage_index, 0, len(cur.stages) - 1))]
    specs = config.joint_specs
    lo, hi = _bounds(specs)
    rng = np.random.default_rng(seed)
    sn, gn = np.asarray(st.start_noise, dtype=float), np.asarray(st.goal_noise, dtype=float)
    any_s, any_g = bool(np.any(sn > 0.0)), bool(np.any(gn > 0.0))
    iq = np.tile(np.asarray(st.start_q, dtype=float), (n_episodes, 1))
    gq = np.tile(np.asarray(st.goal_q, dtype=float), (n_episodes, 1))
    if any_s and any_g:
        # Generator.uniform fills in C order, so one (n,2,7) draw is the same stream as n alternating 7-vector draws
        d = rng.uniform(low=np.stack([-sn, -gn]), high=np.stack([sn, gn]), size=(n_episodes, 2, 7))
        iq, gq = iq + d[:, 0], gq + d[:, 1]
    elif any_g:
        gq = gq + rng.uniform(low=-gn, high=gn, size=(n_episodes, 7))
    elif any_s:
        iq = iq + rng.uniform(low=-sn, high=sn, size=(n_episodes, 7))
    return EvalSuite(initial_q=np.clip(iq, lo, hi), goal_q=np.clip(gq, lo, hi))


def build_fixed_eval_suite(*, seed: int, n_episodes: int, joint_specs: Sequence[JointSpec], start_margin_fraction: float = 0.20,
                           goal_margin_fraction: float = 0.10) -> EvalSuite:
    """``fixed_eval_suite.py:39-63``."""
    rng = np.random.default_rng(seed)
    iq, gq = [], []
    for _ in range(n_episodes):
        iq.append(sample_joint_configuration(rng, joint_specs, start_margin_fraction))
        gq.append(sample_joint_configuration(rng, joint_specs, goal_margin_fraction))
    return EvalSuite(initial_q=np.array(iq).reshape(-1, 7), goal_q=np.array(gq).reshape(-1, 7))


def build_dock_eval_suite(config: Phase1EnvConfig, *, seed: int = 700001, n_episodes: int = 10, fk: FkFn | None = None) -> EvalSuite:
    """``build_dock_eval_suite`` (fixed_eval_suite.py:108-134): ``n_episodes`` draws of ``sample_dock_reset`` at stage 0 from one
    PCG64 stream (no handoff buffer).  ``fk`` is only needed when the config's close-bucket probability is > 0."""
    rng = np.random.default_rng(seed)
    rows = [sample_dock_reset(rng, config, 0, fk) for _ in range(int(n_episodes))]
    stack = lambda name: None if any(getattr(r, name) is None for r in rows) else np.array([getattr(r, name) for r in rows], dtype=float).reshape(len(rows), -1)  # noqa: E731
    return EvalSuite(initial_q=np.array([r.initial_q for r in rows], dtype=float).reshape(-1, 7),
                     goal_q=np.array([r.goal_q for r in rows], dtype=float).reshape(-1, 7), goal_pose6=stack("goal_pose6"),
                     initial_dq=stack("initial_dq"), initial_prev_action=stack("initial_prev_action"))
