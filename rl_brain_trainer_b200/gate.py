"""Workspace-expansion evaluation gate: multi-stage Approach -> Finisher evaluation in one launch + the gate score.

Mirrors ``evaluate_workspace_expansion_checkpoint`` (``kinematic_phase1/eval/eval_workspace_expansion.py:86-211``: per stage a
suite seeded ``seed + stage * 1009``, per-stage summaries) and the gate helpers of
``kinematic_phase1/workspace/workspace_curriculum.py:10-90`` (``stage_passed``, ``retention_ok``, ``highest_passed_stage``,
``gated_score``) that ``WorkspaceEvalGateCallback`` (``train_workspace_expansion.py:54-129``) uses to keep the best checkpoint.
All stages are concatenated into ONE suite and evaluated by one fused-rollout launch; summaries are device reductions.
"""

from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Any, Mapping, Sequence

import numpy as np
import torch

from .config import Phase1EnvConfig
from .policy import PolicyWeights
from .rollout import ApproachFinisherRollout, RolloutResult
from .samplers import EvalSuite, build_curriculum_local_eval_suite


@dataclass(frozen=True)
class WorkspaceGateConfig:
    retention_stage0_4_success: float = 0.95
    retention_stage5_success: float = 0.85
    retention_stage_thresholds: tuple[float, ...] = ()
    promotion_stage_success: float = 0.80
    promotion_ready_rate: float = 0.80
    max_mean_position_error_m: float = 0.020
    max_mean_orientation_error_rad: float = 0.15
    score_current_success_weight: float = 0.45
    score_current_ready_weight: float = 0.20
    score_retention_weight: float = 0.20
    score_error_weight: float = 0.15


def gate_config_from_dict(payload: Mapping[str, Any] | None) -> WorkspaceGateConfig:
    data = dict(payload or {})
    if "retention_stage_thresholds" in data:
        data["retention_stage_thresholds"] = tuple(float(v) for v in data["retention_stage_thresholds"])
    names = {f.name for f in fields(WorkspaceGateConfig)}
    return WorkspaceGateConfig(**{k: v for k, v in data.items() if k in names})


# A stage "passes" when every rule holds: (metric key, value assumed when the key is missing, config bound, +1: at least / -1: at most).
# Known answers of the reference's helpers (workspace_curriculum.py) are pinned in tests/golden/gate_cases.json.
_PASS_RULES = (("success_rate", 0.0, "promotion_stage_success", +1.0),
               ("finisher_ready_hit_rate", 0.0, "promotion_ready_rate", +1.0),
               ("mean_final_position_error", 999.0, "max_mean_position_error_m", -1.0),
               ("mean_final_orientation_error", 999.0, "max_mean_orientation_error_rad", -1.0))


def _metric(row: Mapping[str, Any] | None, key: str, missing: float) -> float:
    return float((row or {}).get(key, missing))


def stage_passed(metrics: Mapping[str, Any], cfg: WorkspaceGateConfig) -> bool:
    return all((_metric(metrics, key, missing) >= getattr(cfg, bound)) if sense > 0 else (_metric(metrics, key, missing) <= getattr(cfg, bound))
               for key, missing, bound, sense in _PASS_RULES)


def _retention_requirements(present: Sequence[int], cfg: WorkspaceGateConfig) -> list[tuple[int, float]]:
    """(stage, minimum success rate) pairs the retention check enforces: an explicit per-stage list binds only the evaluated stages,
    the default binds stages 0-4 and stage 5 whether evaluated or not (a missing stage counts as success 0)."""
    if cfg.retention_stage_thresholds:
        return [(i, float(t)) for i, t in enumerate(cfg.retention_stage_thresholds) if i in present]
    return [(i, cfg.retention_stage0_4_success) for i in range(5)] + [(5, cfg.retention_stage5_success)]


def retention_ok(stage_metrics: Mapping[int, Mapping[str, Any]], cfg: WorkspaceGateConfig) -> bool:
    return all(_metric(stage_metrics.get(i), "success_rate", 0.0) >= need for i, need in _retention_requirements(list(stage_metrics), cfg))


def highest_passed_stage(stage_metrics: Mapping[int, Mapping[str, Any]], cfg: WorkspaceGateConfig) -> int:
    """Largest passing stage index before the first failing frontier stage (index >= 6); -1 if none."""
    order = sorted(stage_metrics)
    verdicts = [stage_passed(stage_metrics[i], cfg) for i in order]
    cut = next((k for k, (i, ok) in enumerate(zip(order, verdicts)) if i >= 6 and not ok), len(order))
    passing = [i for i, ok in zip(order[:cut], verdicts[:cut]) if ok]
    return passing[-1] if passing else -1


def gated_score(stage_metrics: Mapping[int, Mapping[str, Any]], current_stage: int, cfg: WorkspaceGateConfig) -> dict[str, Any]:
    here = stage_metrics.get(current_stage)
    success, ready = _metric(here, "success_rate", 0.0), _metric(here, "finisher_ready_hit_rate", 0.0)
    old_stages = range(0, min(6, current_stage + 1))
    retention = sum(_metric(stage_metrics.get(i), "success_rate", 0.0) for i in old_stages) / len(old_stages) if len(old_stages) else 0.0
    closeness = [max(0.0, 1.0 - _metric(here, key, 1.0) / max(getattr(cfg, bound), 1e-6))
                 for key, bound in (("mean_final_position_error", "max_mean_position_error_m"),
                                    ("mean_final_orientation_error", "max_mean_orientation_error_rad"))]
    error_score = 0.5 * (closeness[0] + closeness[1])
    terms = ((success, cfg.score_current_success_weight), (ready, cfg.score_current_ready_weight), (retention, cfg.score_retention_weight),
             (error_score, cfg.score_error_weight))
    score = 0.0
    for value, weight in terms:
        score += value * weight
    return {"score": float(score), "current_stage": int(current_stage), "retention_ok": retention_ok(stage_metrics, cfg),
            "highest_passed_stage": int(highest_passed_stage(stage_metrics, cfg)), "current_stage_success_rate": success,
            "current_stage_ready_rate": ready, "retention_mean_success_rate": float(retention), "error_score": float(error_score)}


def summarize_stages(result: RolloutResult, stage_of_episode: torch.Tensor, stages: Sequence[int]) -> dict[int, dict[str, float]]:
    """Per-stage ``_summarize_stage`` rows from one concatenated rollout (device segment reductions)."""
    out: dict[int, dict[str, float]] = {}
    n_stage = max(stages) + 1
    idx = stage_of_episode.long()
    count = torch.bincount(idx, minlength=n_stage).double().clamp_min(1)

    def seg(t: torch.Tensor) -> np.ndarray:
        return (torch.bincount(idx, weights=t.double(), minlength=n_stage) / count).cpu().numpy()

    cols = {"success_rate": seg(result.success), "finisher_ready_hit_rate": seg(result.ready_hit), "dwell_success_rate": seg(result.ready_dwell),
            "mean_final_position_error": seg(result.final_position_error), "mean_final_orientation_error": seg(result.final_orientation_error),
            "mean_final_action_magnitude": seg(result.final_action_magnitude), "mean_final_dq_norm": seg(result.final_dq_norm),
            "regression_rate": seg((result.approach_final_position_error > result.min_position_error + 0.002)
                                   | (result.approach_final_orientation_error > result.min_orientation_error + 0.01))}
    counts = torch.bincount(idx, minlength=n_stage).cpu().numpy()
    for s in stages:
        out[int(s)] = {"episode_count": int(counts[s]), **{k: float(v[s]) for k, v in cols.items()}}
    return out


def evaluate_workspace_expansion(approach_config: Phase1EnvConfig, approach_policy: PolicyWeights, finisher_config: Phase1EnvConfig | None,
                                 finisher_policy: PolicyWeights | None, *, episodes: int = 50, seed: int = 700001,
                                 stage_indices: Sequence[int] | None = None, handoff_confirm_steps: int = 2,
                                 gate_config: Mapping[str, Any] | None = None, variant: int = 0, device: str | torch.device = "cuda",
                                 rollout: ApproachFinisherRollout | None = None) -> dict[str, Any]:
    n_stage = len(approach_config.curriculum_config.stages)
    stages = [int(np.clip(s, 0, n_stage - 1)) for s in (stage_indices if stage_indices is not None else range(n_stage))]
    suites = [build_curriculum_local_eval_suite(approach_config, seed=seed + s * 1009, stage_index=s, n_episodes=episodes) for s in stages]
    suite = EvalSuite(initial_q=np.concatenate([x.initial_q for x in suites]), goal_q=np.concatenate([x.goal_q for x in suites]))
    ro = rollout or ApproachFinisherRollout(approach_config, approach_policy, finisher_config, finisher_policy, device=device,
                                            handoff_confirm_steps=handoff_confirm_steps, variant=variant)
    res = ro.evaluate_suite(suite)
    stage_of = torch.as_tensor(np.repeat(np.asarray(stages), episodes), device=res.raw.device)
    metrics = summarize_stages(res, stage_of, stages)
    cfg = gate_config_from_dict(gate_config)
    score_stage = int(np.clip(int((gate_config or {}).get("score_stage_index", max(stages))), min(stages), max(stages)))
    return {"episodes_per_stage": int(episodes), "seed": int(seed), "stage_metrics": metrics,
            "best_model_selection": gated_score(metrics, score_stage, cfg), "env_steps": int(res.env_steps.item())}


class EvalGate:
    """``WorkspaceEvalGateCallback`` for the on-device trainer: every ``eval_interval`` timesteps evaluate, score, keep the best."""

    def __init__(self, approach_config: Phase1EnvConfig, finisher_config: Phase1EnvConfig, finisher_policy: PolicyWeights, *, eval_interval: int,
                 episodes: int, seed: int, stage_indices: Sequence[int], gate_config: Mapping[str, Any], variant: int = 0) -> None:
        self.acfg, self.fcfg, self.fpol = approach_config, finisher_config, finisher_policy
        self.eval_interval, self.episodes, self.seed = max(int(eval_interval), 1), max(int(episodes), 1), int(seed)
        self.stage_indices, self.gate_config, self.variant = list(stage_indices), dict(gate_config), int(variant)
        self.next_eval = self.eval_interval
        self.best_score = float("-inf")
        self.best_state: dict[str, torch.Tensor] | None = None
        self.history: list[dict[str, Any]] = []

    def maybe_eval(self, num_timesteps: int, policy: PolicyWeights, *, force: bool = False) -> dict[str, Any] | None:
        """``force=True`` evaluates regardless of the interval -- e.g. the policy a fine-tune STARTS from, so that the best checkpoint is
        never worse than the initial one (the reference's callback only sees checkpoints after the first interval)."""
        if num_timesteps < self.next_eval and not force:
            return None
        while self.next_eval <= num_timesteps:
            self.next_eval += self.eval_interval
        summary = evaluate_workspace_expansion(self.acfg, policy, self.fcfg, self.fpol, episodes=self.episodes, seed=self.seed,
                                               stage_indices=self.stage_indices, gate_config=self.gate_config, variant=self.variant,
                                               device=policy.device)
        sel = summary["best_model_selection"]
        record = {"timesteps": int(num_timesteps), **sel}
        self.history.append(record)
        if sel["retention_ok"] and sel["score"] > self.best_score:
            self.best_score = float(sel["score"])
            self.best_state = {k: v.detach().clone() for k, v in policy.state_dict().items()}
        return record
