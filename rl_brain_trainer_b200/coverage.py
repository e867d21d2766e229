"""Full-workspace coverage evaluation and adaptive bucket priorities at scale.

Mirrors ``kinematic_phase1/eval/eval_full_workspace_coverage.py:75-308`` (``_summarize``, ``_bucket_metrics``,
``evaluate_full_workspace_coverage``), the bucket ids of ``workspace/workspace_target_map.py:51-121`` and
``workspace/adaptive_frontier_sampler.py`` (``classify_bucket``, ``priority_for_category``, ``update_bucket_priorities``).  The
reference walks Python dict rows one episode at a time; here the episodes of all three splits run in fused rollout launches and the
per-bucket / per-source / per-reason statistics are device segment reductions (``torch.bincount``), so a 10^6-pair sweep costs
milliseconds and can feed the adaptive frontier curriculum every few updates.
"""

from __future__ import annotations

from dataclasses import asdict, dataclass
from typing import Any, Sequence

import numpy as np
import torch

from . import workspace as ws
from .config import Phase1EnvConfig
from .kinematics import fk_pose6_folded
from .policy import PolicyWeights
from .rollout import FAILURE_REASONS, VARIANT_FFMA, ApproachFinisherRollout, RolloutResult, failure_reason_codes


# ------------------------------------------------------------------------------------------------
# buckets (workspace_target_map.py:51-121)
# ------------------------------------------------------------------------------------------------
@dataclass
class TargetBuckets:
    code: np.ndarray          # [T] dense bucket index (order of first appearance)
    ids: list[str]            # bucket_id string of every dense index: "x{}_y{}_z{}_o{}_q{}"
    pose6: np.ndarray         # [T,6] FK(q_target), fp64

    @property
    def count(self) -> int:
        return len(self.ids)


def target_buckets(targets: ws.TargetMap, *, xyz_bins: int = 8, ori_bins: int = 6, q_l2_bins: int = 6) -> TargetBuckets:
    """Bucket id of every target: xyz bins span the map's own bounding box (+-1e-6), orientation = |rpy| / pi, joint = |q| / 4.5."""
    pose = np.array([fk_pose6_folded(q) for q in targets.q]).reshape(-1, 6)
    lo, hi = pose[:, :3].min(axis=0) - 1e-6, pose[:, :3].max(axis=0) + 1e-6
    xyz = np.clip(np.floor((pose[:, :3] - lo) / np.maximum(hi - lo, 1e-9) * xyz_bins), 0, xyz_bins - 1).astype(int)
    ori = np.clip(np.floor(np.linalg.norm(pose[:, 3:], axis=1) / np.pi * ori_bins), 0, ori_bins - 1).astype(int)
    qb = np.clip(np.floor(np.linalg.norm(targets.q, axis=1) / 4.5 * q_l2_bins), 0, q_l2_bins - 1).astype(int)
    ids_per_target = [f"x{a}_y{b}_z{c}_o{o}_q{q}" for (a, b, c), o, q in zip(xyz, ori, qb)]
    index: dict[str, int] = {}
    code = np.array([index.setdefault(s, len(index)) for s in ids_per_target], dtype=np.int64)
    return TargetBuckets(code=code, ids=list(index), pose6=pose)


# ------------------------------------------------------------------------------------------------
# adaptive priorities (adaptive_frontier_sampler.py)
# ------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class BucketPriority:
    bucket_id: str
    success_rate: float
    mean_min_error: float
    mean_final_error: float
    previous_success_rate: float | None
    failure_count: int
    category: str
    sampling_priority: float


_PRIORITY = {"mastered": 0.15, "frontier": 1.00, "hard_but_promising": 0.95, "forgetting_risk": 1.10, "stress": 0.25, "too_hard": 0.05}


def classify_bucket(*, success_rate: float, mean_min_error: float, mean_final_error: float, previous_success_rate: float | None = None) -> str:
    if previous_success_rate is not None and previous_success_rate >= 0.75 and success_rate < previous_success_rate - 0.20:
        return "forgetting_risk"
    if success_rate >= 0.85:
        return "mastered"
    if 0.35 <= success_rate < 0.85:
        return "frontier"
    if success_rate < 0.20 and mean_min_error > 0.025:
        return "too_hard"
    if mean_min_error <= 0.012 and mean_final_error > mean_min_error + 0.006:
        return "hard_but_promising"
    return "stress"


def priority_for_category(category: str) -> float:
    return _PRIORITY.get(category, 0.20)


def update_bucket_priorities(bucket_metrics: dict[str, dict[str, Any]]) -> list[BucketPriority]:
    out = []
    for bucket_id, m in bucket_metrics.items():
        sr = float(m.get("success_rate", 0.0))
        mn = float(m.get("mean_min_position_error", m.get("mean_min_error", 999.0)))
        fin = float(m.get("mean_final_position_error", m.get("mean_final_error", 999.0)))
        prev = m.get("previous_success_rate")
        prev = float(prev) if prev is not None else None
        cat = classify_bucket(success_rate=sr, mean_min_error=mn, mean_final_error=fin, previous_success_rate=prev)
        fails = int(m.get("failure_count", 0))
        out.append(BucketPriority(bucket_id, sr, mn, fin, prev, fails, cat, float(priority_for_category(cat) * (1.0 + min(fails, 20) / 40.0))))
    return sorted(out, key=lambda p: p.sampling_priority, reverse=True)      # stable, like the reference's sorted()


# ------------------------------------------------------------------------------------------------
# device reductions
# ------------------------------------------------------------------------------------------------
def _seg_mean(idx: torch.Tensor, values: torch.Tensor, n: int) -> tuple[np.ndarray, np.ndarray]:
    count = torch.bincount(idx, minlength=n)
    total = torch.bincount(idx, weights=values.double(), minlength=n)
    return (total / count.clamp_min(1).double()).cpu().numpy(), count.cpu().numpy()


def bucket_metrics(result: RolloutResult, bucket_code: torch.Tensor, buckets: TargetBuckets,
                   previous: dict[str, dict[str, Any]] | None = None) -> dict[str, dict[str, Any]]:
    """``_bucket_metrics`` (:110-124) for every bucket that received an episode; ``previous`` (an earlier call's output) fills
    ``previous_success_rate`` so :func:`update_bucket_priorities` can flag forgetting."""
    idx, nb = bucket_code.long(), buckets.count
    sr, count = _seg_mean(idx, result.success, nb)
    fin, _ = _seg_mean(idx, result.final_position_error, nb)
    mn, _ = _seg_mean(idx, result.min_position_error, nb)
    fails = torch.bincount(idx, weights=(~result.success).double(), minlength=nb).cpu().numpy()
    out: dict[str, dict[str, Any]] = {}
    for b in np.nonzero(count)[0]:
        bid = buckets.ids[b]
        out[bid] = {"episode_count": int(count[b]), "success_rate": float(sr[b]), "failure_count": int(round(fails[b])),
                    "mean_final_position_error": float(fin[b]), "mean_min_position_error": float(mn[b])}
        if previous is not None and bid in previous:
            out[bid]["previous_success_rate"] = float(previous[bid]["success_rate"])
    return out


def summarize_split(result: RolloutResult, approach_config: Phase1EnvConfig, *, start_source: torch.Tensor, joint_distance_l2: torch.Tensor,
                    ee_position_distance: torch.Tensor, handoff_confirm_steps: int = 2) -> dict[str, Any]:
    """``_summarize`` (:75-107): rates, mean errors, pair distances, failure reasons, success by start source."""
    f = lambda t: float(t.double().mean().item()) if t.numel() else 0.0  # noqa: E731
    reason = failure_reason_codes(result, approach_config, handoff_confirm_steps)
    counts = torch.bincount(reason, minlength=len(FAILURE_REASONS)).cpu().numpy()
    ok = result.success
    by_src, n_src = _seg_mean(start_source.long(), ok, len(ws.START_SOURCES))
    okd = joint_distance_l2[ok]
    return {
        "episode_count": result.n, "success_rate": f(ok), "ready_rate": f(result.ready_hit), "dwell_success_rate": f(result.ready_dwell),
        "mean_final_position_error": f(result.final_position_error), "mean_final_orientation_error": f(result.final_orientation_error),
        "mean_final_action_magnitude": f(result.final_action_magnitude), "mean_final_dq_norm": f(result.final_dq_norm),
        "average_start_target_joint_distance": f(joint_distance_l2), "average_start_target_ee_distance": f(ee_position_distance),
        "max_successful_joint_l2": float(okd.max().item()) if okd.numel() else 0.0,
        "failure_reason_counts": {FAILURE_REASONS[i]: int(c) for i, c in enumerate(counts) if c},
        "success_by_start_source": {ws.START_SOURCES[s]: {"episode_count": int(n_src[s]), "success_rate": float(by_src[s])}
                                    for s in range(len(ws.START_SOURCES)) if n_src[s]},
    }


def evaluate_full_workspace_coverage(approach_config: Phase1EnvConfig, approach_policy: PolicyWeights, finisher_config: Phase1EnvConfig | None = None,
                                     finisher_policy: PolicyWeights | None = None, *, seed: int = 940001, episodes_per_split: int = 96,
                                     stage_samples_per_stage: int = 96, random_target_samples: int = 384, random_start_samples: int = 384,
                                     pair_count: int = 2048, handoff_confirm_steps: int = 2, variant: int = VARIANT_FFMA,
                                     previous_bucket_metrics: dict[str, dict[str, Any]] | None = None,
                                     device: str | torch.device = "cuda") -> dict[str, Any]:
    """``evaluate_full_workspace_coverage`` (:193-291) without the file writing: maps -> pairs -> the three splits (seed schedule
    seed+1/+2/+3 and the split-selection stream of ``seed``) -> one fused rollout per split -> split summaries, bucket metrics,
    covered / stable / partial / stress bucket fractions and the top sampling priorities."""
    rng = np.random.default_rng(seed)
    targets = ws.generate_workspace_target_map(approach_config, seed=seed + 1, stage_samples_per_stage=stage_samples_per_stage, random_samples=random_target_samples)
    starts = ws.generate_workspace_start_state_map(approach_config, seed=seed + 2, stage_samples_per_stage=max(stage_samples_per_stage // 2, 1),
                                                   random_samples=random_start_samples)
    pairs = ws.build_pair_table(starts, targets, seed=seed + 3, pair_count=pair_count)
    buckets = target_buckets(targets)
    start_pos = np.array([fk_pose6_folded(q)[:3] for q in starts.q]).reshape(-1, 3)
    ro = ApproachFinisherRollout(approach_config, approach_policy, finisher_config, finisher_policy, device=device,
                                 handoff_confirm_steps=handoff_confirm_steps, variant=variant)
    dev = torch.device(device)
    names = {"known": "random_start_known_workspace", "frontier": "random_start_frontier", "stress": "full_reachable_stress"}
    out: dict[str, Any] = {}
    results, codes = [], []
    env_steps = 0
    for split in ("known", "frontier", "stress"):
        sel = ws.select_pairs(pairs, targets, mode=split, limit=episodes_per_split, rng=rng)
        res = ro.evaluate_suite(ws.pairs_to_suite(starts, targets, pairs, sel))
        si, ti = pairs.start[sel], pairs.target[sel]
        ee_dist = np.linalg.norm(buckets.pose6[ti, :3] - start_pos[si], axis=1)
        out[names[split]] = summarize_split(res, approach_config, start_source=torch.as_tensor(starts.source[si], device=dev),
                                            joint_distance_l2=torch.as_tensor(pairs.q_l2[sel], device=dev),
                                            ee_position_distance=torch.as_tensor(ee_dist, device=dev), handoff_confirm_steps=handoff_confirm_steps)
        results.append(res)
        codes.append(torch.as_tensor(buckets.code[ti], device=dev))
        env_steps += int(res.env_steps.item())
    n_all = sum(r.n for r in results)
    merged = RolloutResult(raw=torch.cat([r.raw[:, : r.n] for r in results], dim=1).contiguous(), n=n_all,
                           env_steps=torch.tensor([env_steps], dtype=torch.int64, device=dev))
    metrics = bucket_metrics(merged, torch.cat(codes), buckets, previous_bucket_metrics)
    priorities = update_bucket_priorities(metrics)
    stable = sum(1 for m in metrics.values() if m["success_rate"] >= 0.85)
    partial = sum(1 for m in metrics.values() if 0.35 <= m["success_rate"] < 0.85)
    stress = sum(1 for m in metrics.values() if m["success_rate"] < 0.35)
    nb = max(len(metrics), 1)
    out.update({
        "target_map_summary": {"seed": seed + 1, "total_target_count": int(targets.q.shape[0]), "bucket_count": buckets.count},
        "start_state_map_summary": {"seed": seed + 2, "total_start_count": int(starts.q.shape[0])},
        "pair_sampler_summary": {"seed": seed + 3, "pair_count": int(pairs.start.shape[0]),
                                 "difficulty_class_counts": {ws.DIFFICULTY_CLASSES[k]: int(c) for k, c in enumerate(np.bincount(pairs.klass, minlength=5)) if c}},
        "covered_bucket_fraction": float((stable + partial) / nb), "stable_bucket_fraction": float(stable / nb),
        "partial_bucket_fraction": float(partial / nb), "stress_bucket_fraction": float(stress / nb),
        "covered_bucket_count": int(stable + partial), "total_eval_bucket_count": len(metrics),
        "top_sampling_priorities": [asdict(p) for p in priorities[:30]], "bucket_metrics": metrics, "env_steps": env_steps,
    })
    return out


__all__ = ["BucketPriority", "TargetBuckets", "bucket_metrics", "classify_bucket", "evaluate_full_workspace_coverage", "priority_for_category",
           "summarize_split", "target_buckets", "update_bucket_priorities"]
