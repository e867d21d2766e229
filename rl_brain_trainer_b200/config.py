"""Configuration objects of the kinematic env, mirroring the reference's names and defaults.

The reference resolves YAML -> frozen dataclasses (``kinematic_phase1/training/policy_config.py:72-164``,
``Phase1EnvConfig`` at ``envs/arm_kinematic_env.py:32-66``).  A user switching from the reference keeps
their YAML files: :func:`load_env_config_yaml` applies the same ``base_config`` chaining and deep-merge
(``train_workspace_expansion.py:34-51``) and :func:`to_env_config` the same key lookups and fallbacks,
so the resulting object has the same attribute names and values.  Only the sub-configs the hot path
uses are modelled (approach / dock reward, termination, observation, curriculum, dock reset,
workspace stage sampling); the bridge / dock-coarse stages the reference removed from its final
pipeline are out of scope and their YAML sections are ignored.

Defaults are stated as tables below (name -> default) rather than hand-written classes; the values are
the reference's (``reward_approach.py:13-72``, ``reward_dock.py:13-102``, ``termination.py:11-17``,
``observation_builder.py:18-21``, ``reset_samplers.py:48-64``, ``reward_route.py:14-33``).
"""

from __future__ import annotations

import json
import math
from dataclasses import dataclass, field, make_dataclass, replace
from pathlib import Path
from typing import Any, Iterable, Mapping

N_JOINTS = 7
PRESET_DIR = Path(__file__).resolve().parent / "presets"

JOINT_ORDER = ("Rack_joint", "robot_base_joint", "shoulder1_joint", "shoulder2_joint", "wr1_joint", "wr2_joint", "wr3_joint")


@dataclass(frozen=True)
class JointSpec:
    name: str
    lower: float
    upper: float
    delta_limit: float
    continuous: bool = False

    @property
    def span(self) -> float:
        return float(self.upper - self.lower)


def default_joint_specs() -> tuple[JointSpec, ...]:
    """Hard-coded limits (``kinematics/joint_limits.py:37-47``); the URDF override is absent upstream."""
    deltas = (0.08, 0.30, 0.24, 0.24, 0.30, 0.40, 0.30)
    out = []
    for i, (name, dl) in enumerate(zip(JOINT_ORDER, deltas)):
        lim = 0.385 if i == 0 else math.pi
        out.append(JointSpec(name, -lim, lim, dl, continuous=(name == "wr2_joint")))
    return tuple(out)


def _table_dataclass(name: str, table: Mapping[str, Any], doc: str) -> type:
    fields = []
    for key, default in table.items():
        if isinstance(default, tuple):
            fields.append((key, tuple, field(default=default)))
        else:
            fields.append((key, type(default), field(default=default)))
    cls = make_dataclass(name, fields, frozen=True)
    cls.__doc__ = doc
    cls.__module__ = __name__
    return cls


_APPROACH_REWARD_DEFAULTS: dict[str, Any] = dict(
    position_progress_weight=8.0, orientation_progress_weight=1.0, near_field_orientation_progress_weight=2.0,
    pre_near_goal_pos_threshold_m=0.12, near_goal_pos_threshold_m=0.05, near_goal_ori_threshold_rad=0.35,
    coarse_orientation_bonus_threshold_rad=0.35, orientation_milestone_thresholds_rad=(), orientation_milestone_bonuses=(),
    near_field_orientation_center_weight=0.0, use_orientation_gate=False, pre_near_goal_bonus=0.03, near_goal_bonus=0.10,
    near_goal_bonus_decay=0.5, pre_near_to_near_progress_weight=0.0, coarse_orientation_bonus=0.04,
    handover_pos_threshold_m=0.0, handover_ori_threshold_rad=0.0, handover_bonus=0.0, handover_retention_bonus=0.0,
    handover_dwell_bonus=0.0, handover_leave_penalty=0.0, handover_regression_weight=0.0, handover_smoothness_multiplier=1.0,
    dock_coarse_ready_pos_threshold_m=0.0, dock_coarse_ready_ori_threshold_rad=0.0, dock_coarse_ready_action_threshold=0.0,
    dock_coarse_ready_dq_threshold=0.0, dock_coarse_ready_bonus=0.0, dock_coarse_ready_retention_bonus=0.0,
    dock_coarse_ready_dwell_bonus=0.0, dock_coarse_ready_leave_penalty=0.0, dock_coarse_ready_regression_weight=0.0,
    finisher_ready_pos_threshold_m=0.0, finisher_ready_ori_threshold_rad=0.0, finisher_ready_action_threshold=0.0,
    finisher_ready_dq_threshold=0.0, finisher_ready_bonus=0.0, finisher_ready_retention_bonus=0.0,
    finisher_ready_dwell_bonus=0.0, finisher_ready_leave_penalty=0.0, finisher_ready_regression_weight=0.0,
    near_handoff_pos_threshold_m=0.0, near_handoff_ori_threshold_rad=0.0, near_handoff_action_weight=0.0,
    near_handoff_dq_weight=0.0, near_handoff_motion_bonus_weight=0.0, near_handoff_settle_bonus_weight=0.0,
    same_step_alignment_bonus=0.0, dwell_bonus=0.12, drift_penalty_weight=3.0, drift_penalty_escalation_start=2,
    drift_penalty_escalation_per_count=0.5, near_goal_leave_penalty=0.0, action_magnitude_weight=0.002,
    action_delta_weight=0.004, joint_limit_penalty_weight=0.05, success_bonus=1.0,
)

_DOCK_REWARD_DEFAULTS: dict[str, Any] = dict(
    position_progress_weight=6.0, orientation_progress_weight=5.0, stay_in_zone_bonus=0.08, dwell_bonus=0.18,
    leave_zone_penalty=0.25, working_range_bonus=0.0, working_range_dwell_bonus=0.0, working_range_dwell_start=2,
    working_range_exit_penalty=0.0, drift_penalty_position_weight=4.0, drift_penalty_orientation_weight=2.0,
    action_magnitude_weight=0.006, action_delta_weight=0.012, joint_limit_penalty_weight=0.05, success_bonus=2.0,
    tight_pose_pos_threshold_m=0.005, tight_pose_ori_threshold_rad=0.05, tight_pose_bonus=0.0, tight_pose_dwell_bonus=0.0,
    strict_pose_leave_penalty=0.0, strict_center_reward_weight=0.0, strict_center_position_weight=0.0,
    strict_center_orientation_weight=0.0, strict_center_small_action_bonus_weight=0.0,
    strict_center_small_action_pos_radius_m=0.0, strict_center_small_action_ori_radius_rad=0.0,
    strict_center_small_action_scale=0.0, strict_center_small_action_power=2.0, strict_center_dwell_bonus_weight=0.0,
    strict_center_dwell_start=2, strict_center_dwell_escalation_start=5, strict_center_dwell_escalation_per_step=0.0,
    strict_zone_drift_penalty_multiplier=1.0, strict_zone_action_penalty_multiplier=1.0,
    tight_position_shaping_radius_m=0.0, tight_position_shaping_weight=0.0, tight_orientation_shaping_radius_rad=0.0,
    tight_orientation_shaping_weight=0.0, convergence_position_radius_m=0.0, convergence_position_progress_weight=0.0,
    convergence_orientation_radius_rad=0.0, convergence_orientation_progress_weight=0.0,
    position_first_orientation_pos_threshold_m=0.0, position_first_orientation_pre_scale=1.0,
    action_delta_violation_threshold=0.0, action_delta_violation_weight=0.0, delta_q_change_penalty_threshold=0.0,
    delta_q_change_penalty_weight=0.0, entry_action_penalty_near_pos_threshold_m=0.0,
    entry_action_penalty_far_pos_threshold_m=0.0, entry_action_penalty_near_multiplier=1.0,
    entry_action_penalty_far_multiplier=1.0, basin_outer_radius_m=0.0, basin_inner_radius_m=0.0, basin_dwell_radius_m=0.0,
    basin_outer_bonus=0.0, basin_inner_bonus=0.0, basin_dwell_bonus=0.0, basin_outer_exit_penalty=0.0,
    basin_inner_exit_penalty=0.0, basin_dwell_break_penalty=0.0, basin_drift_penalty_weight=0.0,
    near_strict_pos_threshold_m=0.0, near_strict_ori_threshold_rad=0.0, preserve_state_bonus=0.0,
    preserve_position_tolerance_m=0.0, preserve_orientation_tolerance_rad=0.0, strict_hold_bonus=0.0, low_motion_bonus=0.0,
    low_motion_action_threshold=0.0, low_motion_dq_threshold=0.0, tiny_correction_bonus=0.0,
    tiny_correction_action_threshold=0.0, worse_than_entry_position_weight=0.0, worse_than_entry_orientation_weight=0.0,
    worse_than_entry_position_tolerance_m=0.0, worse_than_entry_orientation_tolerance_rad=0.0,
    near_strict_regression_multiplier=1.0, aggressive_action_weight=0.0, aggressive_action_threshold=0.0,
    dq_penalty_weight=0.0, dq_penalty_threshold=0.0, near_strict_action_penalty_multiplier=1.0,
    near_strict_dq_penalty_multiplier=1.0,
)

_ROUTE_REWARD_DEFAULTS: dict[str, Any] = dict(
    q_goal_progress_weight=2.0, ee_position_progress_weight=6.0, ee_orientation_progress_weight=5.0,
    route_tangent_progress_weight=0.25, same_step_route_ready_bonus=1.5, route_ready_dwell_bonus=0.8,
    low_motion_near_waypoint_bonus=0.4, orientation_regression_penalty_weight=4.0, q_route_regression_penalty_weight=1.0,
    off_route_penalty_weight=0.25, action_magnitude_weight=0.02, action_delta_weight=0.03, dq_penalty_weight=0.8,
    no_progress_penalty=0.02, route_ready_pos_threshold_m=0.010, route_ready_ori_threshold_rad=0.150,
    route_ready_q_threshold=0.080, route_ready_action_threshold=0.25, route_ready_dq_threshold=0.010,
)

_TERMINATION_DEFAULTS: dict[str, Any] = dict(
    max_episode_steps=75, success_pos_threshold_m=0.06, success_ori_threshold_rad=0.15, success_dwell_steps=2,
    require_orientation=False, terminate_on_success=True,
)

_OBSERVATION_DEFAULTS: dict[str, Any] = dict(pos_err_scale_m=0.5, ori_err_scale_rad=math.pi)

_DOCK_RESET_DEFAULTS: dict[str, Any] = dict(
    goal_q=(0.0,) * 7, goal_noise=(0.01, 0.03, 0.04, 0.03, 0.02, 0.02, 0.01),
    init_q_noise=(0.01, 0.02, 0.03, 0.02, 0.015, 0.015, 0.01), close_bucket_probability=0.0,
    close_init_q_noise=(0.006, 0.012, 0.018, 0.012, 0.009, 0.009, 0.006), close_bucket_min_pos_error_m=0.005,
    close_bucket_max_pos_error_m=0.020, close_bucket_min_ori_error_rad=0.0, close_bucket_max_ori_error_rad=0.12,
    close_bucket_max_attempts=128, handoff_state_probability=0.0, handoff_state_buffer_path="",
    handoff_state_max_position_error_m=1.0, handoff_state_max_orientation_error_rad=10.0, handoff_state_max_action_l2=10.0,
)

_ROUTE_RESET_SAMPLER_DEFAULTS: dict[str, Any] = dict(
    mode="mixed_prefix_segment", min_route_index=1, max_route_index=20, segment_start_index=1, segment_end_index=40,
    replay_start_index=1, replay_end_index=120, prefix_start_reset_ratio=0.10, random_prefix_reset_ratio=0.55,
    segment_reset_ratio=0.20, replay_reset_ratio=0.0, recovery_reset_ratio=0.15, q_noise_std=0.002, dq_noise_std=0.0005,
    prev_action_noise_std=0.02,
)

ApproachRewardConfig = _table_dataclass("ApproachRewardConfig", _APPROACH_REWARD_DEFAULTS, "Approach reward knobs (reward_approach.py:13-72).")
DockRewardConfig = _table_dataclass("DockRewardConfig", _DOCK_REWARD_DEFAULTS, "Finisher (dock) reward knobs (reward_dock.py:13-102).")
RouteRewardConfig = _table_dataclass("RouteRewardConfig", _ROUTE_REWARD_DEFAULTS, "Route reward knobs (route/reward_route.py:14-33).")
TerminationConfig = _table_dataclass("TerminationConfig", _TERMINATION_DEFAULTS, "termination.py:11-17.")
ObservationBuilderConfig = _table_dataclass("ObservationBuilderConfig", _OBSERVATION_DEFAULTS, "observation_builder.py:18-21.")
DockResetConfig = _table_dataclass("DockResetConfig", _DOCK_RESET_DEFAULTS, "reset_samplers.py:48-64 (handoff buffer loaded separately).")
RouteResetSamplerConfig = _table_dataclass("RouteResetSamplerConfig", _ROUTE_RESET_SAMPLER_DEFAULTS, "route/route_reset_samplers.py:14-30.")


def _vec7(values: Iterable[float]) -> tuple[float, ...]:
    out = tuple(float(v) for v in values)
    if len(out) != N_JOINTS:
        raise ValueError("Phase 1 curriculum stages require 7-joint vectors")
    return out


@dataclass(frozen=True)
class CurriculumStageConfig:
    """One difficulty shell: start/goal joint centre + uniform noise half-widths (curriculum.py:22-34)."""

    name: str
    start_q: tuple[float, ...]
    goal_q: tuple[float, ...]
    start_noise: tuple[float, ...] = (0.0,) * 7
    goal_noise: tuple[float, ...] = (0.0,) * 7

    def __post_init__(self) -> None:
        for key in ("start_q", "goal_q", "start_noise", "goal_noise"):
            object.__setattr__(self, key, _vec7(getattr(self, key)))


def default_point_curriculum_stages() -> tuple[CurriculumStageConfig, ...]:
    """The 6 default shells (curriculum.py:37-79)."""
    z = (0.0,) * 7
    rows = (
        ("region_small", z, z, (0.01, 0.03, 0.04, 0.03, 0.02, 0.02, 0.01)),
        ("region_medium", z, z, (0.02, 0.06, 0.08, 0.06, 0.04, 0.04, 0.03)),
        ("region_medium_wide", z, z, (0.03, 0.09, 0.12, 0.09, 0.06, 0.05, 0.04)),
        ("region_large", (0.00, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01), z, (0.04, 0.12, 0.16, 0.12, 0.08, 0.06, 0.05)),
        ("region_large_offset", (0.00, 0.02, 0.02, 0.02, 0.02, 0.02, 0.02), (0.03, -0.04, 0.05, -0.03, 0.02, -0.01, 0.01),
         (0.05, 0.14, 0.18, 0.14, 0.09, 0.07, 0.06)),
        ("region_wide_local_random", (0.00, 0.03, 0.03, 0.03, 0.03, 0.03, 0.03), z, (0.06, 0.18, 0.22, 0.16, 0.10, 0.08, 0.07)),
    )
    return tuple(CurriculumStageConfig(name=n, start_q=z, start_noise=sn, goal_q=gq, goal_noise=gn) for n, sn, gq, gn in rows)


@dataclass(frozen=True)
class PointCurriculumConfig:
    enabled: bool = True
    success_rate_threshold: float = 0.80
    window_episodes: int = 20
    min_episodes_per_stage: int = 30
    stages: tuple[CurriculumStageConfig, ...] = field(default_factory=default_point_curriculum_stages)


@dataclass(frozen=True)
class Phase1EnvConfig:
    """Same attribute names as the reference's ``Phase1EnvConfig`` (arm_kinematic_env.py:32-66), hot-path subset."""

    mode_name: str = "approach"
    n_joints: int = 7
    joint_specs: tuple[JointSpec, ...] = field(default_factory=default_joint_specs)
    goal_sample_margin_fraction: float = 0.10
    start_sample_margin_fraction: float = 0.20
    action_delta_scale: float = 1.0
    dynamic_action_delta_scale_enabled: bool = False
    dynamic_action_delta_scale_near_pos_threshold_m: float = 0.0
    dynamic_action_delta_scale_far_pos_threshold_m: float = 0.0
    dynamic_action_delta_scale_near_multiplier: float = 1.0
    dynamic_action_delta_scale_far_multiplier: float = 1.0
    dock_action_delta_scale: float = 0.0
    dock_residual_action_limit: float = 1.0
    dock_delta_q_change_limit_scale: float = 0.0
    dock_dynamic_action_limit_near_pos_threshold_m: float = 0.0
    dock_dynamic_action_limit_far_pos_threshold_m: float = 0.0
    dock_dynamic_residual_action_limit_near: float = 1.0
    dock_dynamic_residual_action_limit_far: float = 1.0
    dock_dynamic_delta_q_change_limit_scale_near: float = 0.0
    dock_dynamic_delta_q_change_limit_scale_far: float = 0.0
    episode_length: int = 75
    dwell_steps_target: int = 3
    curriculum_config: PointCurriculumConfig = field(default_factory=PointCurriculumConfig)
    workspace_stage_sampling: dict = field(default_factory=dict)
    reward_config: Any = field(default_factory=ApproachRewardConfig)
    dock_reward_config: Any = field(default_factory=DockRewardConfig)
    dock_reset_config: Any = field(default_factory=DockResetConfig)
    termination_config: Any = field(default_factory=TerminationConfig)
    observation_config: Any = field(default_factory=ObservationBuilderConfig)


@dataclass(frozen=True)
class RouteObservationConfig:
    include_route_keys: bool = False


@dataclass(frozen=True)
class RouteSequenceConfig:
    """route/route_sequence_env.py:21-26."""

    enabled: bool = False
    sequence_length: int = 5
    reset_ready_streak_on_advance: bool = True


@dataclass(frozen=True)
class RouteEnvConfig:
    """route/route_env.py:19-24."""

    base_env_config: Phase1EnvConfig
    reset_config: Any
    reward_config: Any
    observation_config: RouteObservationConfig = RouteObservationConfig()


# ----------------------------------------------------------------------------------------------
# dict / YAML -> config
# ----------------------------------------------------------------------------------------------
def deep_merge(base: Mapping[str, Any], overlay: Mapping[str, Any]) -> dict[str, Any]:
    """Recursive dict overlay, overlay wins, non-dict values replace (policy_config.py:76-83)."""
    out = dict(base)
    for key, value in overlay.items():
        if isinstance(value, Mapping) and isinstance(out.get(key), Mapping):
            out[key] = deep_merge(out[key], value)
        else:
            out[key] = value
    return out


def _build(cls: type, values: Mapping[str, Any]) -> Any:
    known = {f for f in cls.__dataclass_fields__}
    unknown = set(values) - known
    if unknown:
        # the reference would raise TypeError from the dataclass constructor; keep that behaviour
        raise TypeError(f"{cls.__name__} got unexpected keys {sorted(unknown)}")
    coerced = {}
    for key, value in values.items():
        default = cls.__dataclass_fields__[key].default
        if isinstance(default, tuple) or isinstance(value, list):
            value = tuple(value)
        coerced[key] = value
    return cls(**coerced)


def to_env_config(config: Mapping[str, Any]) -> Phase1EnvConfig:
    """Merged config dict -> :class:`Phase1EnvConfig`; same lookups/fallbacks as policy_config.py:96-164."""
    env = config.get("env", {})
    term = env.get("termination", {})
    cur = env.get("curriculum", {})
    stage_dicts = cur.get("stages")
    if stage_dicts:
        stages = tuple(
            CurriculumStageConfig(name=s["name"], start_q=tuple(s["start_q"]), goal_q=tuple(s["goal_q"]),
                                  start_noise=tuple(s.get("start_noise", [0.0] * 7)), goal_noise=tuple(s.get("goal_noise", [0.0] * 7)))
            for s in stage_dicts)
    else:
        stages = default_point_curriculum_stages()
    f = float
    residual = env.get("dock_residual_action_limit", 1.0)
    dqc = env.get("dock_delta_q_change_limit_scale", 0.0)
    dock_reset = dict(env.get("dock_reset", {}))
    return Phase1EnvConfig(
        mode_name=str(env.get("mode", "approach")),
        n_joints=int(env.get("n_joints", 7)),
        joint_specs=default_joint_specs(),
        goal_sample_margin_fraction=f(env.get("goal_sample_margin_fraction", 0.10)),
        start_sample_margin_fraction=f(env.get("start_sample_margin_fraction", 0.20)),
        action_delta_scale=f(env.get("action_delta_scale", 1.0)),
        dynamic_action_delta_scale_enabled=bool(env.get("dynamic_action_delta_scale_enabled", False)),
        dynamic_action_delta_scale_near_pos_threshold_m=f(env.get("dynamic_action_delta_scale_near_pos_threshold_m", 0.0)),
        dynamic_action_delta_scale_far_pos_threshold_m=f(env.get("dynamic_action_delta_scale_far_pos_threshold_m", 0.0)),
        dynamic_action_delta_scale_near_multiplier=f(env.get("dynamic_action_delta_scale_near_multiplier", 1.0)),
        dynamic_action_delta_scale_far_multiplier=f(env.get("dynamic_action_delta_scale_far_multiplier", 1.0)),
        dock_action_delta_scale=f(env.get("dock_action_delta_scale", 0.0)),
        dock_residual_action_limit=f(residual),
        dock_delta_q_change_limit_scale=f(dqc),
        dock_dynamic_action_limit_near_pos_threshold_m=f(env.get("dock_dynamic_action_limit_near_pos_threshold_m", 0.0)),
        dock_dynamic_action_limit_far_pos_threshold_m=f(env.get("dock_dynamic_action_limit_far_pos_threshold_m", 0.0)),
        dock_dynamic_residual_action_limit_near=f(env.get("dock_dynamic_residual_action_limit_near", residual)),
        dock_dynamic_residual_action_limit_far=f(env.get("dock_dynamic_residual_action_limit_far", residual)),
        dock_dynamic_delta_q_change_limit_scale_near=f(env.get("dock_dynamic_delta_q_change_limit_scale_near", dqc)),
        dock_dynamic_delta_q_change_limit_scale_far=f(env.get("dock_dynamic_delta_q_change_limit_scale_far", dqc)),
        episode_length=int(env.get("episode_length", 75)),
        dwell_steps_target=int(term.get("success_dwell_steps", 3)),  # policy_config.py:146
        curriculum_config=PointCurriculumConfig(
            enabled=bool(cur.get("enabled", True)),
            success_rate_threshold=f(cur.get("success_rate_threshold", 0.80)),
            window_episodes=int(cur.get("window_episodes", 20)),
            min_episodes_per_stage=int(cur.get("min_episodes_per_stage", 30)),
            stages=stages,
        ),
        workspace_stage_sampling=dict(env.get("workspace_stage_sampling", {})),
        reward_config=_build(ApproachRewardConfig, env.get("reward", {})),
        dock_reward_config=_build(DockRewardConfig, env.get("dock_reward", {})),
        dock_reset_config=_build(DockResetConfig, dock_reset) if dock_reset else DockResetConfig(),
        termination_config=_build(TerminationConfig, term),
        observation_config=_build(ObservationBuilderConfig, env.get("observation", {})),
    )


def to_route_env_config(config: Mapping[str, Any], *, max_route_index: int | None = None) -> tuple[RouteEnvConfig, RouteSequenceConfig]:
    """``route:`` section -> wrapper configs (eval_route_curriculum.py:27-47, train_route_curriculum.py)."""
    route = config.get("route", {})
    reset = dict(route.get("reset", {}))
    if max_route_index is not None:
        reset["max_route_index"] = int(max_route_index)
    env_cfg = RouteEnvConfig(
        base_env_config=to_env_config(config),
        reset_config=_build(RouteResetSamplerConfig, reset),
        reward_config=_build(RouteRewardConfig, route.get("reward", {})),
        observation_config=RouteObservationConfig(**route.get("observation", {})),
    )
    return env_cfg, RouteSequenceConfig(**route.get("sequence", {}))


def load_yaml_file(path: str | Path) -> dict[str, Any]:
    import yaml

    return yaml.safe_load(Path(path).read_text()) or {}


def load_overlay_with_bases(path: str | Path, config_dirs: Iterable[Path] = ()) -> dict[str, Any]:
    """``base_config:`` chaining (train_workspace_expansion.py:34-51)."""
    path = Path(path)
    overlay = load_yaml_file(path)
    base = overlay.pop("base_config", None)
    if not base:
        return overlay
    base_path = Path(str(base))
    if not base_path.is_absolute():
        candidates = [path.parent / base_path] + [Path(d) / base_path for d in config_dirs]
        base_path = next((c for c in candidates if c.exists()), candidates[0])
    return deep_merge(load_overlay_with_bases(base_path, config_dirs), overlay)


def load_env_config_yaml(path: str | Path, *, defaults: str | Path | None = None) -> Phase1EnvConfig:
    """Reference-style YAML (+ optional defaults file merged underneath) -> config."""
    merged = load_overlay_with_bases(path)
    if defaults is not None:
        merged = deep_merge(load_yaml_file(defaults), merged)
    return to_env_config(merged)


def preset_dict(name: str) -> dict[str, Any]:
    """Merged config dict of an official run, resolved from the reference's YAMLs by tests/golden/gen_golden.py."""
    path = PRESET_DIR / f"{name}.json"
    if not path.exists():
        raise FileNotFoundError(f"unknown preset {name!r}; have {sorted(p.stem for p in PRESET_DIR.glob('*.json'))}")
    return json.loads(path.read_text())


def load_preset(name: str) -> Phase1EnvConfig:
    return to_env_config(preset_dict(name))


__all__ = [
    "ApproachRewardConfig", "CurriculumStageConfig", "DockResetConfig", "DockRewardConfig", "JointSpec", "N_JOINTS",
    "ObservationBuilderConfig", "Phase1EnvConfig", "PointCurriculumConfig", "RouteEnvConfig", "RouteObservationConfig",
    "RouteResetSamplerConfig", "RouteRewardConfig", "RouteSequenceConfig", "TerminationConfig", "deep_merge",
    "default_joint_specs", "default_point_curriculum_stages", "load_env_config_yaml", "load_overlay_with_bases",
    "load_preset", "load_yaml_file", "preset_dict", "replace", "to_env_config", "to_route_env_config",
]
