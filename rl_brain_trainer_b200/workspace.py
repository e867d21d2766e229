"""Mixed random-start workspace maps and start/target pair tables (host-side bookkeeping).

Mirrors ``kinematic_phase1/workspace/workspace_target_map.py:76-157``, ``workspace_start_state_map.py:62-134``,
``start_target_pair_sampler.py:30-115`` and the split filters of ``eval/eval_full_workspace_coverage.py:58-72``,
consuming the numpy PCG64 stream in the reference's order so a seed reproduces the reference's maps
(checked against ``tests/golden/randomstart_maps.npz``).  Everything is array-valued (no per-row dicts) so the
1M-pair configuration of BASELINE.json is built in well under a second; the per-sample bucket ids and difficulty
scores the reference also stores are report-only metadata and are not part of the rollout path.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .config import Phase1EnvConfig
from .samplers import EvalSuite, sample_joint_configuration, sample_stage_joint_target

START_SOURCES = ("home", "successful_rollout", "near_target", "random_valid_q")
DIFFICULTY_CLASSES = ("retention", "local", "medium", "frontier", "stress")


@dataclass
class TargetMap:
    q: np.ndarray        # [T,7]
    stage: np.ndarray    # [T] int, -1 for uniform-valid targets


@dataclass
class StartMap:
    q: np.ndarray             # [S,7]
    dq: np.ndarray            # [S,7]
    prev_action: np.ndarray   # [S,7]
    source: np.ndarray        # [S] index into START_SOURCES


@dataclass
class PairTable:
    start: np.ndarray        # [P] index into StartMap
    target: np.ndarray       # [P] index into TargetMap
    q_l2: np.ndarray         # [P]
    klass: np.ndarray        # [P] index into DIFFICULTY_CLASSES


def generate_workspace_target_map(config: Phase1EnvConfig, *, seed: int, stage_samples_per_stage: int, random_samples: int) -> TargetMap:
    rng = np.random.default_rng(seed)
    stages = config.curriculum_config.stages
    qs, st = [], []
    for stage_id, stage in enumerate(stages):
        for _ in range(max(stage_samples_per_stage, 0)):
            qs.append(sample_stage_joint_target(rng, stage.goal_q, stage.goal_noise, config.joint_specs))
            st.append(stage_id)
    for _ in range(max(random_samples, 0)):
        qs.append(sample_joint_configuration(rng, config.joint_specs, margin_fraction=0.08))
        st.append(-1)
    return TargetMap(q=np.array(qs).reshape(-1, 7), stage=np.array(st, dtype=np.int64))


def generate_workspace_start_state_map(config: Phase1EnvConfig, *, seed: int, stage_samples_per_stage: int, random_samples: int,
                                       dq_noise: float = 0.001, prev_action_noise: float = 0.03) -> StartMap:
    rng = np.random.default_rng(seed)
    stages = config.curriculum_config.stages
    qs, src = [np.zeros(7)], [0]
    for stage_id, stage in enumerate(stages):
        for _ in range(max(stage_samples_per_stage, 0)):
            if rng.random() < 0.65:
                qs.append(sample_stage_joint_target(rng, stage.goal_q, stage.goal_noise, config.joint_specs))
                src.append(1)
            else:
                qs.append(sample_stage_joint_target(rng, stage.start_q, stage.start_noise, config.joint_specs))
                src.append(2 if stage_id >= 6 else 1)
    for _ in range(max(random_samples, 0)):
        qs.append(sample_joint_configuration(rng, config.joint_specs, margin_fraction=0.10))
        src.append(3)
    n = len(qs)
    # per sample: 7 dq draws then 7 prev-action draws (also for "home", whose draws are then discarded)
    noise = rng.uniform(low=np.array([[-dq_noise], [-prev_action_noise]]), high=np.array([[dq_noise], [prev_action_noise]]), size=(n, 2, 7))
    dq, pa = noise[:, 0].copy(), noise[:, 1].copy()
    dq[0] = 0.0
    pa[0] = 0.0
    return StartMap(q=np.array(qs).reshape(-1, 7), dq=dq, prev_action=pa, source=np.array(src, dtype=np.int64))


def classify_pairs(starts: StartMap, targets: TargetMap, si: np.ndarray, ti: np.ndarray, q_l2: np.ndarray, *, local_q_l2: float = 0.28,
                   medium_q_l2: float = 0.70) -> np.ndarray:
    """``classify_pair`` (start_target_pair_sampler.py:30-50) without eval history (previous_eval_success_rate is None)."""
    tstage = targets.stage[ti]
    ssrc = starts.source[si]
    retention = np.isin(ssrc, (0, 1)) & (tstage >= 0) & (tstage <= 7)
    tstage0 = np.where(tstage < 0, 0, tstage)  # int(target.get("stage_id") or 0)
    klass = np.where(tstage0 <= 10, 3, 4)
    klass = np.where(q_l2 <= medium_q_l2, 2, klass)
    klass = np.where(q_l2 <= local_q_l2, 1, klass)
    return np.where(retention, 0, klass).astype(np.int64)


def build_pair_table(starts: StartMap, targets: TargetMap, *, seed: int, pair_count: int) -> PairTable:
    rng = np.random.default_rng(seed)
    n = max(int(pair_count), 0)
    # the reference draws start index then target index per pair; a C-ordered (n,2) bounded draw is the same stream
    idx = rng.integers(0, np.array([starts.q.shape[0], targets.q.shape[0]]), size=(n, 2))
    si, ti = idx[:, 0], idx[:, 1]
    q_l2 = np.linalg.norm(targets.q[ti] - starts.q[si], axis=1)
    return PairTable(start=si, target=ti, q_l2=q_l2, klass=classify_pairs(starts, targets, si, ti, q_l2))


def select_pairs(pairs: PairTable, targets: TargetMap, *, mode: str, limit: int, rng: np.random.Generator) -> np.ndarray:
    """Indices into the pair table for one split (eval_full_workspace_coverage.py:58-72)."""
    tstage0 = np.where(targets.stage[pairs.target] < 0, 0, targets.stage[pairs.target])
    if mode == "known":
        pool = np.nonzero((tstage0 <= 8) & np.isin(pairs.klass, (0, 1, 2)))[0]
    elif mode == "frontier":
        pool = np.nonzero((tstage0 >= 8) & (tstage0 <= 11) & np.isin(pairs.klass, (2, 3, 4)))[0]
    elif mode == "stress":
        pool = np.arange(pairs.start.shape[0])
    else:
        raise ValueError(f"Unknown pair eval mode: {mode}")
    if pool.size == 0:
        pool = np.arange(pairs.start.shape[0])
    if pool.size <= limit:
        return pool
    return pool[rng.choice(pool.size, size=limit, replace=False)]


def pairs_to_suite(starts: StartMap, targets: TargetMap, pairs: PairTable, sel: np.ndarray) -> EvalSuite:
    si, ti = pairs.start[sel], pairs.target[sel]
    return EvalSuite(initial_q=starts.q[si], goal_q=targets.q[ti], initial_dq=starts.dq[si], initial_prev_action=starts.prev_action[si])


def build_randomstart_eval(config: Phase1EnvConfig, *, seed: int = 940001, episodes_per_split: int = 96, stage_samples_per_stage: int = 96,
                           random_target_samples: int = 384, random_start_samples: int = 384, pair_count: int = 2048) -> dict[str, EvalSuite]:
    """The three splits of ``evaluate_full_workspace_coverage`` (:193-256) with its seed schedule (seed+1/+2/+3)."""
    rng = np.random.default_rng(seed)
    targets = generate_workspace_target_map(config, seed=seed + 1, stage_samples_per_stage=stage_samples_per_stage, random_samples=random_target_samples)
    starts = generate_workspace_start_state_map(config, seed=seed + 2, stage_samples_per_stage=max(stage_samples_per_stage // 2, 1),
                                                random_samples=random_start_samples)
    pairs = build_pair_table(starts, targets, seed=seed + 3, pair_count=pair_count)
    out = {}
    for split in ("known", "frontier", "stress"):
        sel = select_pairs(pairs, targets, mode=split, limit=episodes_per_split, rng=rng)
        out[split] = pairs_to_suite(starts, targets, pairs, sel)
    return out
