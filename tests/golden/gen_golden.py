#!/usr/bin/env python
"""Generate golden fixtures by RUNNING THE LIVE REFERENCE (jerry102102102/RL_brain_trainer).

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden.py

It imports the reference's pure-Python kinematic env from /root/reference, drives it with seeded
inputs and records inputs + outputs.  Nothing here is copied from the reference: the files written
are *outputs* of the reference (golden vectors), the merged YAML configs it resolves (presets),
and the fp32 weights of the four SB3 PPO checkpoints bundled in
report/final_codes_docker_submission.zip.

Outputs
  tests/golden/fk.npz                      FK(q) -> pose6 / 4x4, wrap_to_pi samples
  tests/golden/trace_approach.npz          open-loop step traces, official approach config
  tests/golden/trace_dock.npz              open-loop step traces, finisher (dock) config
  tests/golden/trace_route.npz             RouteKinematicEnv / RouteSequenceKinematicEnv traces (synthetic route)
  tests/golden/eval_stage5.npz             Approach->Finisher eval, stage 5 suite (seed 700001+5*1009), 64 episodes
  tests/golden/eval_stages.npz             16 episodes of stages 0, 8, 11
  tests/golden/eval_randomstart.npz        mixed random-start known/frontier/stress splits (seed 940001, 96 each)
  tests/golden/samplers.npz                seeded reset-sampler draws (numpy PCG64 order)
  rl_brain_trainer_b200/presets/*.json     merged config dicts (input of to_env_config)
  rl_brain_trainer_b200/presets/policies/*.npz   policy weights (SB3 key names)
"""

from __future__ import annotations

import dataclasses
import io
import json
import sys
import tempfile
import zipfile
from pathlib import Path

import numpy as np
import torch

REF_ROOT = Path("/root/reference")
REF_PKG = REF_ROOT / "hrl_ws/src/hrl_trainer"
sys.dont_write_bytecode = True
sys.path.insert(0, str(REF_PKG))

REPO = Path(__file__).resolve().parents[2]
GOLD = REPO / "tests" / "golden"
PRESETS = REPO / "rl_brain_trainer_b200" / "presets"

from hrl_trainer.kinematic_phase1.envs.arm_kinematic_env import ArmKinematicEnv  # noqa: E402
from hrl_trainer.kinematic_phase1.envs import reset_samplers as ref_samplers  # noqa: E402
from hrl_trainer.kinematic_phase1.eval import eval_full_workspace_coverage as ref_cov  # noqa: E402
from hrl_trainer.kinematic_phase1.eval.eval_approach_finisher import _finisher_ready  # noqa: E402
from hrl_trainer.kinematic_phase1.eval.eval_pipeline_ablation import _run_approach_with_handoff  # noqa: E402
from hrl_trainer.kinematic_phase1.eval.eval_three_stage import _run_policy, _state_reset_options  # noqa: E402
from hrl_trainer.kinematic_phase1.eval.fixed_eval_suite import build_curriculum_local_eval_suite  # noqa: E402
from hrl_trainer.kinematic_phase1.kinematics.fk_interface import compute_ee_pose6  # noqa: E402
from hrl_trainer.kinematic_phase1.kinematics.pose_utils import wrap_to_pi  # noqa: E402
from hrl_trainer.kinematic_phase1.route.reward_route import RouteRewardConfig  # noqa: E402
from hrl_trainer.kinematic_phase1.route.route_dataset import load_route_dataset  # noqa: E402
from hrl_trainer.kinematic_phase1.route.route_env import RouteEnvConfig, RouteKinematicEnv  # noqa: E402
from hrl_trainer.kinematic_phase1.route.route_observation import RouteObservationConfig  # noqa: E402
from hrl_trainer.kinematic_phase1.route.route_reset_samplers import RouteResetSamplerConfig  # noqa: E402
from hrl_trainer.kinematic_phase1.route.route_sequence_env import RouteSequenceConfig, RouteSequenceKinematicEnv  # noqa: E402
from hrl_trainer.kinematic_phase1.training.policy_config import (  # noqa: E402
    approach_default_config_path,
    config_dir,
    deep_merge,
    load_yaml_file,
    to_env_config,
)
from hrl_trainer.kinematic_phase1.eval.eval_workspace_expansion import _load_overlay_with_bases  # noqa: E402
from hrl_trainer.v5_1.ee_fk import fk_matrix_from_q7  # noqa: E402

ZIP = REF_ROOT / "report/final_codes_docker_submission.zip"
CKPT = {
    "approach_stage8_11": "artifacts/kinematic_phase1/workspace_expansion/workspace_expand_dynscale_stage8_11_big_001/best_checkpoint/model_best_by_gate.zip",
    "finisher": "artifacts/kinematic_phase1/phase1c/dock_workspace_handoff_noop_ft_1m_001/model_latest.zip",
    "randomstart": "artifacts/kinematic_phase1/workspace_full_coverage_randomstart/workspace_full_coverage_randomstart_overnight_003/best_checkpoint/model_best_by_gate.zip",
    "route_prefix120": "artifacts/kinematic_phase1/route_curriculum/route_prefix120_routeobs_sequence2_1m_001/model_latest.zip",
}
OBS_KEYS_56 = sorted(["q", "dq", "prev_action", "goal_pos_err", "goal_ori_err", "wp_pos_err", "wp_ori_err",
                      "next_wp_pos_err", "next_wp_ori_err", "task_type", "mode_flag", "progress", "joint_limit_margin"])
OBS_KEYS_80 = sorted(OBS_KEYS_56 + ["route_q_goal", "route_q_error", "route_tangent", "route_scalar"])


# ------------------------------------------------------------------------------------------
# policies: plain-torch stand-in for SB3 `model.predict(obs, deterministic=True)`
# ------------------------------------------------------------------------------------------
class TorchPolicy:
    def __init__(self, sd: dict[str, torch.Tensor]) -> None:
        self.sd = {k: v.float() for k, v in sd.items()}
        self.in_dim = int(self.sd["mlp_extractor.policy_net.0.weight"].shape[1])
        self.keys = OBS_KEYS_56 if self.in_dim == 56 else OBS_KEYS_80

    def flat(self, obs: dict[str, np.ndarray]) -> np.ndarray:
        return np.concatenate([np.asarray(obs[k], dtype=np.float32).reshape(-1) for k in self.keys])

    def predict(self, obs, deterministic: bool = True):
        x = torch.from_numpy(self.flat(obs))
        sd = self.sd
        with torch.no_grad():
            h = torch.tanh(torch.nn.functional.linear(x, sd["mlp_extractor.policy_net.0.weight"], sd["mlp_extractor.policy_net.0.bias"]))
            h = torch.tanh(torch.nn.functional.linear(h, sd["mlp_extractor.policy_net.2.weight"], sd["mlp_extractor.policy_net.2.bias"]))
            a = torch.nn.functional.linear(h, sd["action_net.weight"], sd["action_net.bias"])
        return np.clip(a.numpy(), -1.0, 1.0), None

    def value(self, obs) -> float:
        x = torch.from_numpy(self.flat(obs))
        sd = self.sd
        with torch.no_grad():
            h = torch.tanh(torch.nn.functional.linear(x, sd["mlp_extractor.value_net.0.weight"], sd["mlp_extractor.value_net.0.bias"]))
            h = torch.tanh(torch.nn.functional.linear(h, sd["mlp_extractor.value_net.2.weight"], sd["mlp_extractor.value_net.2.bias"]))
            v = torch.nn.functional.linear(h, sd["value_net.weight"], sd["value_net.bias"])
        return float(v[0])


def load_checkpoints() -> tuple[dict[str, TorchPolicy], dict[str, dict]]:
    z = zipfile.ZipFile(ZIP)
    policies, hyper = {}, {}
    for name, member in CKPT.items():
        inner = zipfile.ZipFile(io.BytesIO(z.read(member)))
        sd = torch.load(io.BytesIO(inner.read("policy.pth")), weights_only=True, map_location="cpu")
        policies[name] = TorchPolicy(sd)
        data = json.loads(inner.read("data"))
        hyper[name] = {k: data[k] for k in ("learning_rate", "n_steps", "batch_size", "n_epochs", "gamma", "gae_lambda",
                                            "ent_coef", "vf_coef", "max_grad_norm", "normalize_advantage", "n_envs")
                       if k in data and not isinstance(data[k], dict)}
        cr = data.get("clip_range")
        hyper[name]["clip_range_repr"] = str(cr)[:200] if cr is not None else None
    return policies, hyper


# ------------------------------------------------------------------------------------------
# configs
# ------------------------------------------------------------------------------------------
def merged_configs() -> dict[str, dict]:
    cdir = config_dir()
    base = load_yaml_file(approach_default_config_path())
    approach = deep_merge(base, _load_overlay_with_bases(cdir / "workspace_expansion_dynamic_scale_big.yaml"))
    randomstart = deep_merge(base, _load_overlay_with_bases(cdir / "workspace_full_coverage_randomstart_overnight.yaml"))
    finisher = load_yaml_file(cdir / "dock_workspace_handoff_noop_ft_12env.yaml")
    # the handoff-state buffer artifact is not shipped: blank the path so the config loads
    finisher["env"]["dock_reset"]["handoff_state_buffer_path"] = ""
    route = deep_merge(base, _load_overlay_with_bases(cdir / "route_curriculum_prefix120_routeobs_sequence2.yaml"))
    default = dict(base)
    return {"approach_dynamic_scale_big": approach, "randomstart_overnight": randomstart, "finisher_noop_ft": finisher,
            "route_prefix120": route, "approach_default": default}


def route_reward_cfg(cfg: dict) -> RouteRewardConfig:
    return RouteRewardConfig(**cfg.get("route", {}).get("reward", {}))


# ------------------------------------------------------------------------------------------
# recording helpers
# ------------------------------------------------------------------------------------------
def flat56(obs: dict[str, np.ndarray]) -> np.ndarray:
    return np.concatenate([np.asarray(obs[k], dtype=np.float32).reshape(-1) for k in OBS_KEYS_56])


def flat80(obs: dict[str, np.ndarray]) -> np.ndarray:
    return np.concatenate([np.asarray(obs[k], dtype=np.float32).reshape(-1) for k in OBS_KEYS_80])


REASON_CODE = {"running": 0, "success": 1, "max_steps": 2, "invalid_state": 3, "reset": 0}


class TraceRecorder:
    def __init__(self) -> None:
        self.rows: dict[str, list] = {}
        self.episode_start: list[int] = []
        self.resets: dict[str, list] = {}
        self.n = 0

    def reset(self, env: ArmKinematicEnv, options: dict, obs, info, mode_index: int) -> None:
        self.episode_start.append(self.n)
        r = self.resets
        r.setdefault("initial_q", []).append(np.asarray(options["initial_q"], dtype=float))
        r.setdefault("initial_dq", []).append(np.asarray(options.get("initial_dq", np.zeros(7)), dtype=float))
        r.setdefault("initial_prev_action", []).append(np.asarray(options.get("initial_prev_action", np.zeros(7)), dtype=float))
        r.setdefault("goal_q", []).append(np.asarray(options.get("goal_q", np.zeros(7)), dtype=float))
        r.setdefault("goal_pose6", []).append(np.asarray(info["goal_pose6"], dtype=float))
        r.setdefault("has_goal_pose6", []).append(int("goal_pose6" in options))
        r.setdefault("mode", []).append(mode_index)
        r.setdefault("reset_obs", []).append(flat56(obs))
        r.setdefault("reset_ee_pose6", []).append(np.asarray(info["ee_pose6"], dtype=float))
        r.setdefault("entry", []).append(np.array([info["entry_position_error_norm"], info["entry_orientation_error_norm"],
                                                   info["entry_action_l2"], info["entry_dq_norm"]]))

    def step(self, action, obs, reward, terminated, truncated, info, env: ArmKinematicEnv) -> None:
        rows = self.rows
        comps = np.array(list(info["reward_components"].values()), dtype=float)
        comps = np.pad(comps, (0, 64 - len(comps)))
        rows.setdefault("action", []).append(np.asarray(action, dtype=float))
        rows.setdefault("obs", []).append(flat56(obs))
        rows.setdefault("reward", []).append(float(reward))
        rows.setdefault("components", []).append(comps)
        rows.setdefault("terminated", []).append(int(terminated))
        rows.setdefault("truncated", []).append(int(truncated))
        rows.setdefault("success", []).append(int(info["success"]))
        rows.setdefault("reason", []).append(REASON_CODE.get(info["reason"], 9))
        rows.setdefault("q", []).append(np.asarray(info["q"], dtype=float))
        rows.setdefault("dq", []).append(np.asarray(info["dq"], dtype=float))
        rows.setdefault("prev_action", []).append(np.asarray(env._prev_action, dtype=float))
        rows.setdefault("ee_pose6", []).append(np.asarray(info["ee_pose6"], dtype=float))
        rows.setdefault("pos_err", []).append(float(info["position_error_norm"]))
        rows.setdefault("ori_err", []).append(float(info["orientation_error_norm"]))
        rows.setdefault("action_l2", []).append(float(info["action_l2"]))
        rows.setdefault("dq_l2", []).append(float(info["executed_delta_q_l2"]))
        rows.setdefault("dq_change_l2", []).append(float(info["delta_q_change_l2"]))
        rows.setdefault("dock_action_limit", []).append(float(info["dock_action_limit"]))
        rows.setdefault("min_pos_err", []).append(float(info["min_position_error"]))
        rows.setdefault("margin_min", []).append(float(info["joint_limit_margin_min"]))
        rows.setdefault("counters", []).append(np.array([info["step_count"], info["dwell_count"], info["near_goal_entry_count"],
                                                         info["near_goal_drift_count"], int(info["pre_near_goal_hit"]),
                                                         int(info["near_goal_hit"]), int(info["curr_in_pre_near_goal"]),
                                                         int(info["curr_in_near_goal"])], dtype=np.int64))
        self.n += 1

    def arrays(self) -> dict[str, np.ndarray]:
        out = {k: np.asarray(v) for k, v in self.rows.items()}
        out.update({"reset_" + k if not k.startswith("reset_") else k: np.asarray(v) for k, v in self.resets.items()})
        out["episode_start"] = np.asarray(self.episode_start + [self.n], dtype=np.int64)
        return out


def run_trace(rec: TraceRecorder, env: ArmKinematicEnv, options: dict, mode: str, action_fn, max_steps: int) -> dict:
    opts = {**options, "policy_mode": mode}
    obs, info = env.reset(options=opts)
    rec.reset(env, opts, obs, info, 0 if mode == "approach" else 1)
    for t in range(max_steps):
        a = action_fn(obs, t)
        obs, reward, terminated, truncated, info = env.step(a)
        rec.step(a, obs, reward, terminated, truncated, info, env)
        if terminated or truncated:
            break
    return info


# ------------------------------------------------------------------------------------------
def gen_fk() -> None:
    rng = np.random.default_rng(20260101)
    q = rng.uniform(-np.pi, np.pi, size=(512, 7))
    q[:, 0] = rng.uniform(-0.385, 0.385, size=512)
    q[0] = 0.0
    q[1] = np.array([0.12, -0.35, 0.48, -0.62, 0.27, -0.14, 0.51])  # the reference's own FK probe point (TESTS/test_v5_1_ee_fk_external_consistency.py)
    q[2:130] *= 0.15  # stage-shell sized configurations
    pose = np.array([compute_ee_pose6(x) for x in q])
    mats = np.array([fk_matrix_from_q7(x) for x in q[:32]])
    w_in = np.concatenate([rng.uniform(-12, 12, size=200), np.array([np.pi, -np.pi, 0.0, 3 * np.pi, -3 * np.pi, 2 * np.pi])])
    w_out = np.array([float(wrap_to_pi(v)) for v in w_in])
    np.savez_compressed(GOLD / "fk.npz", q=q, pose6=pose, mats=mats, wrap_in=w_in, wrap_out=w_out)
    print("fk.npz", q.shape)


def gen_traces(cfgs, policies) -> dict[str, list[dict]]:
    rng = np.random.default_rng(20260102)
    approach_cfg = to_env_config(cfgs["approach_dynamic_scale_big"])
    finisher_cfg = to_env_config(cfgs["finisher_noop_ft"])
    pol_a, pol_f = policies["approach_stage8_11"], policies["finisher"]
    handoffs: list[dict] = []

    rec = TraceRecorder()
    stages = approach_cfg.curriculum_config.stages
    # (1) policy-driven episodes on several stage shells (exercise near-goal/handoff reward terms)
    for stage_index in (0, 3, 5, 8, 11):
        suite = build_curriculum_local_eval_suite(approach_cfg, seed=1234 + stage_index, stage_index=stage_index, n_episodes=2)
        for ep in suite:
            env = ArmKinematicEnv(approach_cfg)
            env.set_curriculum_stage(stage_index)
            info = run_trace(rec, env, ep.reset_options(), "approach", lambda o, t: pol_a.predict(o)[0].astype(float), 128)
            handoffs.append({"initial_q": info["q"], "initial_dq": info["dq"], "initial_prev_action": env._prev_action.copy(),
                             "goal_q": info["goal_q"], "goal_pose6": info["goal_pose6"]})
    # (2) policy + noise (leave/re-enter zones, drift counters)
    for stage_index in (2, 5):
        suite = build_curriculum_local_eval_suite(approach_cfg, seed=4321 + stage_index, stage_index=stage_index, n_episodes=1)
        env = ArmKinematicEnv(approach_cfg)
        noise = rng.normal(0.0, 0.08, size=(128, 7))
        run_trace(rec, env, suite[0].reset_options(), "approach",
                  lambda o, t: pol_a.predict(o)[0].astype(float) + (noise[t] if (t // 16) % 2 else 0.0), 128)
    # (3) random out-of-range actions from random valid starts (clip paths, joint limits)
    for _ in range(2):
        q0 = rng.uniform(-3.0, 3.0, size=7); q0[0] = rng.uniform(-0.38, 0.38)
        g = rng.uniform(-2.5, 2.5, size=7); g[0] = rng.uniform(-0.3, 0.3)
        acts = rng.uniform(-1.6, 1.6, size=(128, 7))
        env = ArmKinematicEnv(approach_cfg)
        run_trace(rec, env, {"initial_q": q0, "goal_q": g, "initial_dq": rng.uniform(-0.01, 0.01, 7),
                             "initial_prev_action": rng.uniform(-0.5, 0.5, 7)}, "approach", lambda o, t: acts[t], 128)
    # (4) start AT the goal with tiny actions (success/dwell path; reference test_kinematic_phase1_env.py:51-60)
    g = np.asarray(stages[4].goal_q, dtype=float)
    acts = rng.normal(0.0, 0.01, size=(40, 7)); acts[:6] = 0.0
    env = ArmKinematicEnv(approach_cfg)
    run_trace(rec, env, {"initial_q": g, "goal_q": g}, "approach", lambda o, t: acts[t], 40)
    # (5) saturate a joint limit
    q0 = np.array([0.37, 3.0, -3.0, 0.0, 0.0, 3.1, 0.0])
    acts = np.tile(np.array([1.0, 1.0, -1.0, 0.3, -0.3, 1.0, 0.0]), (24, 1))
    env = ArmKinematicEnv(approach_cfg)
    run_trace(rec, env, {"initial_q": q0, "goal_q": np.zeros(7)}, "approach", lambda o, t: acts[t], 24)
    np.savez_compressed(GOLD / "trace_approach.npz", **rec.arrays())
    print("trace_approach.npz steps", rec.n)

    rec = TraceRecorder()
    # dock traces from real handoff states, policy-driven and noisy
    for i, h in enumerate(handoffs):
        env = ArmKinematicEnv(finisher_cfg)
        if i % 2 == 0:
            run_trace(rec, env, h, "dock", lambda o, t: pol_f.predict(o)[0].astype(float), 36)
        else:
            noise = rng.normal(0.0, 0.35, size=(36, 7))
            run_trace(rec, env, h, "dock", lambda o, t: pol_f.predict(o)[0].astype(float) + noise[t], 36)
    # random actions near a goal (dynamic limits, dq rate limit if configured, basin terms off)
    for _ in range(3):
        g = rng.uniform(-0.5, 0.5, size=7); g[0] = rng.uniform(-0.1, 0.1)
        q0 = g + rng.uniform(-0.004, 0.004, size=7)
        acts = rng.uniform(-1.3, 1.3, size=(36, 7)) * rng.choice([0.02, 0.2, 1.0])
        env = ArmKinematicEnv(finisher_cfg)
        run_trace(rec, env, {"initial_q": q0, "goal_q": g, "initial_dq": rng.uniform(-0.002, 0.002, 7),
                             "initial_prev_action": rng.uniform(-0.1, 0.1, 7)}, "dock", lambda o, t: acts[t], 36)
    np.savez_compressed(GOLD / "trace_dock.npz", **rec.arrays())
    print("trace_dock.npz steps", rec.n)

    # a dock config that switches ON the knobs the official finisher leaves off
    # (dynamic action limit + delta-q rate limit + basin shaping + working range), so those branches are pinned too
    alt = json.loads(json.dumps(cfgs["finisher_noop_ft"]))
    alt["env"].update({"dock_action_delta_scale": 0.02, "dock_residual_action_limit": 0.6, "dock_delta_q_change_limit_scale": 0.5,
                       "dock_dynamic_action_limit_near_pos_threshold_m": 0.004, "dock_dynamic_action_limit_far_pos_threshold_m": 0.02,
                       "dock_dynamic_residual_action_limit_near": 0.11, "dock_dynamic_residual_action_limit_far": 0.9,
                       "dock_dynamic_delta_q_change_limit_scale_near": 0.04, "dock_dynamic_delta_q_change_limit_scale_far": 0.8})
    alt["env"]["dock_reward"].update({"basin_outer_radius_m": 0.02, "basin_inner_radius_m": 0.01, "basin_dwell_radius_m": 0.005,
                                      "basin_outer_bonus": 0.1, "basin_inner_bonus": 0.2, "basin_dwell_bonus": 0.3,
                                      "basin_outer_exit_penalty": 0.4, "basin_inner_exit_penalty": 0.5,
                                      "basin_dwell_break_penalty": 0.6, "basin_drift_penalty_weight": 2.0,
                                      "working_range_bonus": 0.05, "working_range_dwell_bonus": 0.02, "working_range_exit_penalty": 0.3,
                                      "stay_in_zone_bonus": 0.08, "dwell_bonus": 0.18, "leave_zone_penalty": 0.25,
                                      "entry_action_penalty_near_pos_threshold_m": 0.003, "entry_action_penalty_far_pos_threshold_m": 0.012,
                                      "entry_action_penalty_near_multiplier": 0.5, "entry_action_penalty_far_multiplier": 2.0})
    alt["env"]["termination"]["terminate_on_success"] = True
    alt_cfg = to_env_config(alt)
    rec = TraceRecorder()
    for i in range(4):
        g = rng.uniform(-0.6, 0.6, size=7); g[0] = rng.uniform(-0.1, 0.1)
        q0 = g + rng.uniform(-0.006, 0.006, size=7)
        acts = rng.uniform(-1.3, 1.3, size=(36, 7)) * (0.05 if i < 2 else 0.6)
        if i == 0:
            q0 = g.copy(); acts[:12] = 0.0
        env = ArmKinematicEnv(alt_cfg)
        run_trace(rec, env, {"initial_q": q0, "goal_q": g, "initial_dq": rng.uniform(-0.001, 0.001, 7) * (i > 0),
                             "initial_prev_action": rng.uniform(-0.05, 0.05, 7) * (i > 0)}, "dock", lambda o, t: acts[t], 36)
    np.savez_compressed(GOLD / "trace_dock_alt.npz", **rec.arrays())
    (GOLD / "trace_dock_alt_config.json").write_text(json.dumps(alt, indent=1))
    print("trace_dock_alt.npz steps", rec.n)

    # an approach config with the handover / milestone families switched on
    alt = json.loads(json.dumps(cfgs["approach_dynamic_scale_big"]))
    alt["env"]["reward"].update({"orientation_milestone_thresholds_rad": [0.3, 0.15, 0.05], "orientation_milestone_bonuses": [0.01, 0.02, 0.03],
                                 "near_field_orientation_center_weight": 0.2, "pre_near_to_near_progress_weight": 1.5,
                                 "handover_pos_threshold_m": 0.01, "handover_ori_threshold_rad": 0.1, "handover_bonus": 0.5,
                                 "handover_retention_bonus": 0.1, "handover_dwell_bonus": 0.2, "handover_leave_penalty": 0.7,
                                 "handover_regression_weight": 3.0, "handover_smoothness_multiplier": 2.5,
                                 "dock_coarse_ready_bonus": 0.3, "dock_coarse_ready_retention_bonus": 0.2,
                                 "dock_coarse_ready_dwell_bonus": 0.1, "dock_coarse_ready_leave_penalty": 0.4,
                                 "dock_coarse_ready_regression_weight": 2.0, "near_goal_leave_penalty": 0.35, "dwell_bonus": 0.12})
    alt["env"]["termination"]["terminate_on_success"] = True
    alt["env"]["termination"]["success_dwell_steps"] = 12
    alt["env"]["dynamic_action_delta_scale_enabled"] = False
    alt_cfg = to_env_config(alt)
    rec = TraceRecorder()
    for i, stage_index in enumerate((1, 4, 6)):
        suite = build_curriculum_local_eval_suite(alt_cfg, seed=99 + stage_index, stage_index=stage_index, n_episodes=1)
        noise = rng.normal(0.0, 0.05, size=(128, 7))
        env = ArmKinematicEnv(alt_cfg)
        run_trace(rec, env, suite[0].reset_options(), "approach",
                  lambda o, t: pol_a.predict(o)[0].astype(float) + (noise[t] if t > 40 and (t // 8) % 2 else 0.0), 128)
    np.savez_compressed(GOLD / "trace_approach_alt.npz", **rec.arrays())
    (GOLD / "trace_approach_alt_config.json").write_text(json.dumps(alt, indent=1))
    print("trace_approach_alt.npz steps", rec.n)
    return {"handoffs": handoffs}


def eval_pipeline(approach_cfg, finisher_cfg, pol_a, pol_f, reset_options_list, stage_index=None) -> dict[str, np.ndarray]:
    """eval_workspace_expansion.py:126-173 / eval_full_workspace_coverage.py:120-190 inner pipeline."""
    rows: dict[str, list] = {}
    for opts in reset_options_list:
        env = ArmKinematicEnv(config=approach_cfg)
        if stage_index is not None:
            env.set_curriculum_stage(stage_index)
        env.set_policy_mode("approach")
        approach_result, handoff_result = _run_approach_with_handoff(
            env=env, model=pol_a, reset_options={**opts, "policy_mode": "approach"}, ready_cfg=approach_cfg.reward_config,
            handoff_confirm_steps=2)
        final_ready = _finisher_ready(approach_result, cfg=approach_cfg.reward_config)
        first_handoff = handoff_result
        handoff_result = approach_result if final_ready else handoff_result
        final_result = approach_result
        success = bool(approach_result["success"])
        if handoff_result is not None:
            fenv = ArmKinematicEnv(config=finisher_cfg)
            fenv.set_policy_mode("dock")
            final_result = _run_policy(env=fenv, model=pol_f, reset_options=_state_reset_options(handoff_result, policy_mode="dock"))
            success = bool(final_result["success"])
        r = rows
        r.setdefault("success", []).append(int(success))
        r.setdefault("approach_success", []).append(int(approach_result["success"]))
        r.setdefault("final_position_error", []).append(float(final_result["final_position_error"]))
        r.setdefault("final_orientation_error", []).append(float(final_result["final_orientation_error"]))
        r.setdefault("approach_final_position_error", []).append(float(approach_result["final_position_error"]))
        r.setdefault("approach_final_orientation_error", []).append(float(approach_result["final_orientation_error"]))
        r.setdefault("min_position_error", []).append(float(approach_result["min_position_error"]))
        r.setdefault("min_orientation_error", []).append(float(approach_result["min_orientation_error"]))
        r.setdefault("final_action_magnitude", []).append(float(final_result["final_action_magnitude"]))
        r.setdefault("final_dq_norm", []).append(float(final_result["final_dq_norm"]))
        r.setdefault("ready_hit", []).append(int(approach_result["dock_coarse_ready_hit"] or final_ready))
        r.setdefault("ready_dwell", []).append(int(approach_result["dock_coarse_ready_dwell"] or final_ready))
        r.setdefault("final_ready", []).append(int(final_ready))
        r.setdefault("first_ready_step", []).append(-1 if approach_result["first_dock_coarse_ready_step"] is None else int(approach_result["first_dock_coarse_ready_step"]))
        r.setdefault("max_ready_streak", []).append(int(approach_result["max_dock_coarse_ready_streak"]))
        r.setdefault("handoff_kind", []).append(2 if final_ready else (1 if first_handoff is not None else 0))
        r.setdefault("handoff_step", []).append(int(handoff_result["step_count"]) if handoff_result is not None else -1)
        r.setdefault("approach_steps", []).append(int(approach_result["step_count"]))
        r.setdefault("finisher_steps", []).append(int(final_result["step_count"]) if handoff_result is not None else 0)
        r.setdefault("final_q", []).append(np.asarray(final_result["final_q"], dtype=float))
        r.setdefault("initial_q", []).append(np.asarray(opts["initial_q"], dtype=float))
        r.setdefault("initial_dq", []).append(np.asarray(opts.get("initial_dq", np.zeros(7)), dtype=float))
        r.setdefault("initial_prev_action", []).append(np.asarray(opts.get("initial_prev_action", np.zeros(7)), dtype=float))
        r.setdefault("goal_q", []).append(np.asarray(opts["goal_q"], dtype=float))
        r.setdefault("goal_pose6", []).append(np.asarray(opts["goal_pose6"], dtype=float))
    return {k: np.asarray(v) for k, v in rows.items()}


def gen_eval(cfgs, policies) -> None:
    approach_cfg = to_env_config(cfgs["approach_dynamic_scale_big"])
    finisher_cfg = to_env_config(cfgs["finisher_noop_ft"])
    pol_a, pol_f = policies["approach_stage8_11"], policies["finisher"]
    seed = 700001
    suite = build_curriculum_local_eval_suite(approach_cfg, seed=seed + 5 * 1009, stage_index=5, n_episodes=64)
    res = eval_pipeline(approach_cfg, finisher_cfg, pol_a, pol_f, [ep.reset_options() for ep in suite], stage_index=5)
    np.savez_compressed(GOLD / "eval_stage5.npz", suite_seed=seed + 5 * 1009, stage_index=5, **res)
    print("eval_stage5.npz success", res["success"].mean(), "pos", res["final_position_error"].mean(), "ori", res["final_orientation_error"].mean())
    out = {}
    for stage_index in (0, 8, 11):
        suite = build_curriculum_local_eval_suite(approach_cfg, seed=seed + stage_index * 1009, stage_index=stage_index, n_episodes=16)
        res = eval_pipeline(approach_cfg, finisher_cfg, pol_a, pol_f, [ep.reset_options() for ep in suite], stage_index=stage_index)
        print("stage", stage_index, "success", res["success"].mean())
        for k, v in res.items():
            out[f"s{stage_index}_{k}"] = v
    np.savez_compressed(GOLD / "eval_stages.npz", **out)


def gen_randomstart(cfgs, policies) -> None:
    """eval_full_workspace_coverage.py:193-256 with its defaults (seed 940001)."""
    cdir = config_dir()
    cfg_path = cdir / "workspace_full_coverage_randomstart_overnight.yaml"
    approach_cfg = to_env_config(cfgs["randomstart_overnight"])
    finisher_cfg = to_env_config(cfgs["finisher_noop_ft"])
    pol_a, pol_f = policies["randomstart"], policies["finisher"]
    seed = 940001
    rng = np.random.default_rng(seed)
    targets, _ = ref_cov.generate_workspace_target_map(config_path=cfg_path, seed=seed + 1, stage_samples_per_stage=96, random_samples=384)
    starts, _ = ref_cov.generate_workspace_start_state_map(config_path=cfg_path, seed=seed + 2, stage_samples_per_stage=48, random_samples=384)
    with tempfile.TemporaryDirectory() as td:
        map_dir = Path(td)
        ref_cov.write_target_map(targets, {}, map_dir)
        ref_cov.write_start_state_map(starts, {}, map_dir)
        pairs, _ = ref_cov.build_pair_sampler_summary(start_map_path=map_dir / "start_state_map.jsonl",
                                                      target_map_path=map_dir / "target_map.jsonl", seed=seed + 3, pair_count=2048)
        starts_by_id = {row["start_id"]: row for row in ref_cov.load_jsonl(map_dir / "start_state_map.jsonl")}
        targets_by_id = {row["target_id"]: row for row in ref_cov.load_jsonl(map_dir / "target_map.jsonl")}
    # the maps + pair table themselves (host-side sampler parity)
    src_types = sorted({s.source_type for s in starts})
    np.savez_compressed(
        GOLD / "randomstart_maps.npz",
        target_q=np.array([t.q_target for t in targets]), target_stage=np.array([-1 if t.stage_id is None else t.stage_id for t in targets]),
        target_pose6=np.array([[*t.ee_target_position, *t.ee_target_orientation] for t in targets]),
        start_q=np.array([s.q_start for s in starts]), start_dq=np.array([s.dq_start for s in starts]),
        start_prev_action=np.array([s.prev_action for s in starts]),
        start_source=np.array([src_types.index(s.source_type) for s in starts]), start_source_names=np.array(src_types),
        pair_start=np.array([int(p["start_id"].split("_")[1]) for p in pairs]), pair_target=np.array([int(p["target_id"].split("_")[1]) for p in pairs]),
        pair_class=np.array([["retention", "local", "medium", "frontier", "stress"].index(p["difficulty_class"]) for p in pairs]),
        pair_q_l2=np.array([p["joint_distance_l2"] for p in pairs]),
    )
    out = {}
    for split in ("known", "frontier", "stress"):
        sel = ref_cov._select_pairs(pairs, mode=split, limit=96, rng=rng)
        opts = []
        for pair in sel:
            s, t = starts_by_id[pair["start_id"]], targets_by_id[pair["target_id"]]
            opts.append({"initial_q": s["q_start"], "initial_dq": s.get("dq_start", [0.0] * 7), "initial_prev_action": s.get("prev_action", [0.0] * 7),
                         "goal_q": t["q_target"], "goal_pose6": [*t["ee_target_position"], *t["ee_target_orientation"]]})
        res = eval_pipeline(approach_cfg, finisher_cfg, pol_a, pol_f, opts)
        print("randomstart", split, "success", res["success"].mean(), "n", len(sel))
        out[f"{split}_pair_index"] = np.array([int(p["pair_id"].split("_")[1]) for p in sel])
        for k, v in res.items():
            out[f"{split}_{k}"] = v
    np.savez_compressed(GOLD / "eval_randomstart.npz", seed=seed, **out)


def synthetic_route(n: int = 40, seed: int = 7) -> np.ndarray:
    """A small dense q-route (the real 483-waypoint file is not in the snapshot)."""
    rng = np.random.default_rng(seed)
    knots = np.cumsum(rng.normal(0.0, 0.22, size=(7, 7)), axis=0) * np.array([0.2, 1, 1, 1, 1, 1, 1])
    t = np.linspace(0, len(knots) - 1, n)
    i0 = np.clip(np.floor(t).astype(int), 0, len(knots) - 2)
    w = (t - i0)[:, None]
    w = w * w * (3 - 2 * w)
    return knots[i0] * (1 - w) + knots[i0 + 1] * w


def gen_route(cfgs, policies) -> None:
    cfg = cfgs["route_prefix120"]
    base_cfg = to_env_config(cfg)
    rr = route_reward_cfg(cfg)
    route_q = synthetic_route()
    with tempfile.TemporaryDirectory() as td:
        path = Path(td) / "route.json"
        path.write_text(json.dumps({"route_q": route_q.tolist()}))
        route = load_route_dataset(path)
    env_cfg = RouteEnvConfig(base_env_config=base_cfg, reset_config=RouteResetSamplerConfig(max_route_index=len(route) - 1),
                             reward_config=rr, observation_config=RouteObservationConfig(include_route_keys=True))
    pol = policies["route_prefix120"]
    rng = np.random.default_rng(5)
    rows: dict[str, list] = {}

    def rec(prefix, action, obs, reward, terminated, truncated, info):
        r = rows
        r.setdefault(prefix + "action", []).append(np.asarray(action, dtype=float))
        r.setdefault(prefix + "obs", []).append(flat80(obs))
        r.setdefault(prefix + "reward", []).append(float(reward))
        r.setdefault(prefix + "components", []).append(np.array(list(info["reward_components"].values()), dtype=float))
        r.setdefault(prefix + "flags", []).append(np.array([int(terminated), int(truncated), int(info["success"]), int(info["route_ready"]),
                                                              int(info["route_ready_streak"]), int(info["route_regression"]),
                                                              int(info["route_orientation_hit"]), int(info["route_index"])]))
        r.setdefault(prefix + "scalars", []).append(np.array([info["route_q_error_norm"], info["nearest_route_q_distance"],
                                                                info["position_error_norm"], info["orientation_error_norm"]]))
        r.setdefault(prefix + "q", []).append(np.asarray(info["q"], dtype=float))

    # RouteKinematicEnv: sequential chain over the first 12 waypoints with the route policy (+ noise on odd ones)
    env = RouteKinematicEnv(route=route, config=env_cfg)
    cq, cdq, cpa = route.waypoint(0).q_goal.copy(), np.zeros(7), np.zeros(7)
    ep_start = [0]
    n = 0
    for idx in range(1, 13):
        env.reset(options={"route_index": idx, "start_route_index": 0, "policy_mode": "approach"})
        obs, info = env.base_env.reset(options={"initial_q": cq, "initial_dq": cdq, "initial_prev_action": cpa,
                                                "goal_q": route.waypoint(idx).q_goal, "policy_mode": "approach"})
        env._route_index, env._start_route_index, env._ready_streak, env._prev_info = idx, 0, 0, dict(info)
        obs = env._augment_obs(obs)
        rows.setdefault("seq_reset_obs", []).append(flat80(obs))
        rows.setdefault("seq_reset_q", []).append(cq.copy()); rows.setdefault("seq_reset_dq", []).append(cdq.copy())
        rows.setdefault("seq_reset_pa", []).append(cpa.copy()); rows.setdefault("seq_reset_index", []).append(idx)
        terminated = truncated = False
        while not (terminated or truncated):
            a = pol.predict(obs)[0].astype(float)
            if idx % 3 == 0:
                a = np.clip(a + rng.normal(0, 0.3, 7), -1, 1)
            obs, reward, terminated, truncated, info = env.step(a)
            rec("seq_", a, obs, reward, terminated, truncated, info)
            n += 1
        ep_start.append(n)
        cq, cdq, cpa = np.asarray(info["q"], float), np.asarray(info["dq"], float), env.base_env._prev_action.copy()
    rows["seq_episode_start"] = ep_start

    # RouteSequenceKinematicEnv: in-episode waypoint advance (sequence_length 4)
    senv = RouteSequenceKinematicEnv(route=route, config=env_cfg, sequence_config=RouteSequenceConfig(enabled=True, sequence_length=4))
    ep_start = [0]
    n = 0
    for first in (1, 12, 30):
        obs, info = senv.reset(options={"route_index": first, "start_route_index": first - 1})
        rows.setdefault("adv_reset_obs", []).append(flat80(obs)); rows.setdefault("adv_reset_index", []).append(first)
        terminated = truncated = False
        while not (terminated or truncated):
            a = pol.predict(obs)[0].astype(float)
            obs, reward, terminated, truncated, info = senv.step(a)
            rec("adv_", a, obs, reward, terminated, truncated, info)
            rows.setdefault("adv_completed", []).append(int(info["route_completed_waypoints"]))
            n += 1
        ep_start.append(n)
    rows["adv_episode_start"] = ep_start
    np.savez_compressed(GOLD / "trace_route.npz", route_q=route_q, route_pose6=route.poses6,
                        route_progress=np.array([wp.route_progress_m for wp in route.waypoints]),
                        **{k: np.asarray(v) for k, v in rows.items()})
    print("trace_route.npz seq steps", len(rows["seq_reward"]), "adv steps", len(rows["adv_reward"]),
          "seq successes", int(np.sum([rows["seq_flags"][i - 1][2] for i in rows["seq_episode_start"][1:]])))


def gen_samplers(cfgs) -> None:
    """Seeded draws of the reference reset samplers (numpy PCG64 call order is part of the contract)."""
    out = {}
    approach_cfg = to_env_config(cfgs["approach_dynamic_scale_big"])
    rs_cfg = to_env_config(cfgs["randomstart_overnight"])
    fin_cfg = to_env_config(cfgs["finisher_noop_ft"])
    for name, cfg, stage_list in (("approach", approach_cfg, (0, 5, 11)), ("randomstart", rs_cfg, (8, 11))):
        for stage in stage_list:
            env = ArmKinematicEnv(cfg)
            env.set_curriculum_stage(stage)
            qs, gs, gps, dqs, pas = [], [], [], [], []
            env.reset(seed=1000 + stage)
            for _ in range(64):
                _, info = env.reset()
                qs.append(info["q"]); gs.append(info["goal_q"]); gps.append(info["goal_pose6"]); dqs.append(info["dq"]); pas.append(env._prev_action.copy())
            out[f"{name}_s{stage}_q"] = np.array(qs); out[f"{name}_s{stage}_goal_q"] = np.array(gs)
            out[f"{name}_s{stage}_goal_pose6"] = np.array(gps); out[f"{name}_s{stage}_dq"] = np.array(dqs)
            out[f"{name}_s{stage}_prev_action"] = np.array(pas)
    env = ArmKinematicEnv(fin_cfg)
    env.reset(seed=77)
    qs, gs = [], []
    for _ in range(64):
        _, info = env.reset()
        qs.append(info["q"]); gs.append(info["goal_q"])
    out["dock_q"] = np.array(qs); out["dock_goal_q"] = np.array(gs)
    # stage-5 eval suite episodes beyond the 64 used by eval_stage5 (suite builder parity)
    suite = build_curriculum_local_eval_suite(approach_cfg, seed=700001 + 5 * 1009, stage_index=5, n_episodes=256)
    out["suite5_initial_q"] = np.array([e.initial_q for e in suite]); out["suite5_goal_q"] = np.array([e.goal_q for e in suite])
    out["suite5_goal_pose6"] = np.array([e.goal_pose6 for e in suite])
    np.savez_compressed(GOLD / "samplers.npz", **out)
    print("samplers.npz", len(out))
    assert ref_samplers is not None


def gen_gate() -> None:
    """Gate-score known answers from the reference's workspace_curriculum.py on seeded random stage metrics."""
    from hrl_trainer.kinematic_phase1.workspace.workspace_curriculum import gate_config_from_dict, gated_score

    rng = np.random.default_rng(12)
    cases = []
    gate_cfgs = [None, {"score_stage_index": 9, "retention_stage0_4_success": 0.95, "retention_stage5_success": 0.85, "promotion_stage_success": 0.55,
                        "promotion_ready_rate": 0.62, "max_mean_position_error_m": 0.024, "max_mean_orientation_error_rad": 0.16},
                 {"retention_stage_thresholds": [0.98, 0.98, 0.98, 0.95, 0.90, 0.90, 0.85, 0.75], "promotion_stage_success": 0.55, "promotion_ready_rate": 0.60}]
    for i in range(24):
        gc = gate_cfgs[i % 3]
        n = int(rng.integers(6, 13))
        hi = i % 4 == 0
        metrics = {s: {"success_rate": float(np.clip(rng.normal(0.99 if hi else 0.9 - 0.05 * s, 0.02 if hi else 0.1), 0, 1)),
                       "finisher_ready_hit_rate": float(np.clip(rng.normal(0.9 - 0.03 * s, 0.1), 0, 1)),
                       "mean_final_position_error": float(abs(rng.normal(0.003 + 0.002 * s, 0.002))),
                       "mean_final_orientation_error": float(abs(rng.normal(0.02 + 0.01 * s, 0.01)))} for s in range(n)}
        cur = int(rng.integers(0, n))
        out = gated_score(metrics, cur, gate_config_from_dict(gc))
        cases.append({"gate_config": gc, "stage_metrics": {str(k): v for k, v in metrics.items()}, "current_stage": cur, "expected": out})
    (GOLD / "gate_cases.json").write_text(json.dumps(cases, indent=1))
    print("gate_cases.json", len(cases))


def write_presets(cfgs, policies, hyper) -> None:
    PRESETS.mkdir(parents=True, exist_ok=True)
    (PRESETS / "policies").mkdir(exist_ok=True)
    for name, cfg in cfgs.items():
        (PRESETS / f"{name}.json").write_text(json.dumps(cfg, indent=1, sort_keys=True))
    for name, pol in policies.items():
        np.savez_compressed(PRESETS / "policies" / f"{name}.npz", **{k: v.numpy() for k, v in pol.sd.items()})
    (PRESETS / "policies" / "hyperparams.json").write_text(json.dumps(hyper, indent=1, sort_keys=True))
    print("presets written")


def main() -> None:
    GOLD.mkdir(parents=True, exist_ok=True)
    policies, hyper = load_checkpoints()
    cfgs = merged_configs()
    write_presets(cfgs, policies, hyper)
    gen_fk()
    gen_gate()
    gen_traces(cfgs, policies)
    gen_samplers(cfgs)
    gen_route(cfgs, policies)
    gen_eval(cfgs, policies)
    gen_randomstart(cfgs, policies)
    assert dataclasses is not None


if __name__ == "__main__":
    main()
