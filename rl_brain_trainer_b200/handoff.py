"""Handoff-state buffer for Finisher adaptation, built by GPU Approach rollouts.

Replaces ``kinematic_phase1/training/build_finisher_handoff_state_buffer.py:44-147`` (the builder) and
``envs/reset_samplers.py:131-166`` (``_load_handoff_states``: the JSON reader with its error / action filters).  The buffer the
reference's bundled Finisher was trained from is absent upstream; this regenerates it from any Approach checkpoint:
one ``kin_rollout_handoff_states`` launch runs every episode of the suite with the handoff bookkeeping of
``_run_approach_with_handoff`` and returns the approach end state and the first-confirmed snapshot per episode.  The JSON written
has the reference's keys, so the reference's ``DockResetConfig(handoff_state_buffer_path=...)`` reads it; on the device the rows
feed the dock reset's handoff-state replay (``BatchedArmKinematicEnv.set_handoff_states``).
"""

from __future__ import annotations

import ctypes
import json
from dataclasses import dataclass
from pathlib import Path
from typing import Any

import numpy as np
import torch

from . import _lib
from .config import Phase1EnvConfig
from .env import ParamsHandle
from .policy import PolicyWeights
from .samplers import EvalSuite, build_curriculum_local_eval_suite, build_fixed_eval_suite

_D = _lib.define
HANDOFF_MODES = ("final_settled", "first_confirmed", "final_always")


@dataclass
class HandoffBuffer:
    """Column arrays of the stored handoff states (one row per state, the reference's ``states`` list stacked)."""

    episode_id: np.ndarray            # [M] int
    step_index: np.ndarray            # [M] int
    initial_q: np.ndarray             # [M,7]
    initial_dq: np.ndarray
    initial_prev_action: np.ndarray
    goal_q: np.ndarray
    goal_pose6: np.ndarray            # [M,6]
    position_error_norm: np.ndarray   # [M]
    orientation_error_norm: np.ndarray
    action_l2: np.ndarray
    dq_norm: np.ndarray

    def __len__(self) -> int:
        return int(self.initial_q.shape[0])

    def select(self, mask: np.ndarray) -> "HandoffBuffer":
        return HandoffBuffer(**{k: getattr(self, k)[mask] for k in self.__dataclass_fields__})

    def filtered(self, *, max_position_error_m: float = 1.0, max_orientation_error_rad: float = 10.0, max_action_l2: float = 10.0) -> "HandoffBuffer":
        """``_load_handoff_states``' filters (``DockResetConfig.handoff_state_max_*``, reset_samplers.py:146-156)."""
        keep = (self.position_error_norm <= max_position_error_m) & (self.orientation_error_norm <= max_orientation_error_rad) & \
               (self.action_l2 <= max_action_l2)
        return self.select(keep)

    def rows(self) -> np.ndarray:
        """``[M, 34]`` float32: initial_q | initial_dq | initial_prev_action | goal_q | goal_pose6 (``KIN_HANDOFF_STATE_FLOATS``)."""
        return np.concatenate([self.initial_q, self.initial_dq, self.initial_prev_action, self.goal_q, self.goal_pose6], axis=1).astype(np.float32)

    def device_rows(self, device: str | torch.device = "cuda") -> torch.Tensor:
        return torch.as_tensor(self.rows(), device=device).contiguous()

    def to_states(self, *, dwell_count: int, source_checkpoint_name: str = "", handoff_mode: str = "final_settled") -> list[dict[str, Any]]:
        """The reference's ``states`` entries (build_finisher_handoff_state_buffer.py:88-104)."""
        return [{"episode_id": int(self.episode_id[i]), "step_index": int(self.step_index[i]), "initial_q": self.initial_q[i].tolist(),
                 "initial_dq": self.initial_dq[i].tolist(), "initial_prev_action": self.initial_prev_action[i].tolist(),
                 "goal_q": self.goal_q[i].tolist(), "goal_pose6": self.goal_pose6[i].tolist(),
                 "position_error_norm": float(self.position_error_norm[i]), "orientation_error_norm": float(self.orientation_error_norm[i]),
                 "dwell_count": int(dwell_count), "action_l2": float(self.action_l2[i]), "dq_norm": float(self.dq_norm[i]),
                 "source_checkpoint_name": source_checkpoint_name, "handoff_mode": handoff_mode} for i in range(len(self))]

    @classmethod
    def from_states(cls, states: list[dict[str, Any]]) -> "HandoffBuffer":
        f = lambda k, w, d=0.0: np.asarray([s.get(k, [d] * w) for s in states], dtype=np.float64).reshape(len(states), w)  # noqa: E731
        g = lambda k: np.asarray([float(s.get(k, 0.0)) for s in states], dtype=np.float64)  # noqa: E731
        return cls(episode_id=np.asarray([int(s.get("episode_id", i)) for i, s in enumerate(states)]),
                   step_index=np.asarray([int(s.get("step_index", 0)) for s in states]), initial_q=f("initial_q", 7), initial_dq=f("initial_dq", 7),
                   initial_prev_action=f("initial_prev_action", 7), goal_q=f("goal_q", 7), goal_pose6=f("goal_pose6", 6),
                   position_error_norm=g("position_error_norm"), orientation_error_norm=g("orientation_error_norm"), action_l2=g("action_l2"),
                   dq_norm=g("dq_norm"))


def load_handoff_states(path: str | Path, *, max_position_error_m: float = 1.0, max_orientation_error_rad: float = 10.0,
                        max_action_l2: float = 10.0) -> HandoffBuffer:
    """``_load_handoff_states`` (reset_samplers.py:131-166): a ``{"states": [...]}`` payload or a bare list; ``FileNotFoundError`` like the reference."""
    p = Path(path)
    if not p.exists():
        raise FileNotFoundError(f"Handoff state buffer does not exist: {p}")
    payload = json.loads(p.read_text())
    raw = payload.get("states", []) if isinstance(payload, dict) else (payload if isinstance(payload, list) else [])
    return HandoffBuffer.from_states(list(raw)).filtered(max_position_error_m=max_position_error_m,
                                                         max_orientation_error_rad=max_orientation_error_rad, max_action_l2=max_action_l2)


def _finisher_ready(cfg: Any, pos: np.ndarray, ori: np.ndarray, action: np.ndarray, dq: np.ndarray) -> np.ndarray:
    """build_finisher_handoff_state_buffer.py:19-28, vectorised."""
    if not (cfg.finisher_ready_pos_threshold_m > 0.0 and cfg.finisher_ready_ori_threshold_rad > 0.0):
        return np.zeros(pos.shape, dtype=bool)
    ok = (pos <= cfg.finisher_ready_pos_threshold_m) & (ori <= cfg.finisher_ready_ori_threshold_rad)
    if cfg.finisher_ready_action_threshold > 0.0:
        ok &= action <= cfg.finisher_ready_action_threshold
    if cfg.finisher_ready_dq_threshold > 0.0:
        ok &= dq <= cfg.finisher_ready_dq_threshold
    return ok


def build_finisher_handoff_state_buffer(approach_config: Phase1EnvConfig, approach_policy: PolicyWeights, *, episodes: int = 500, seed: int = 700001,
                                        stage_index: int = 0, handoff_confirm_steps: int = 2, handoff_mode: str = "final_settled",
                                        suite: EvalSuite | None = None, device: str | torch.device = "cuda",
                                        source_checkpoint_name: str = "") -> tuple[HandoffBuffer, dict[str, Any]]:
    """The reference builder's ``main`` (:44-143) for ``episodes`` episodes in one launch.  Returns the buffer and the summary dict
    (``states`` and ``episode_summaries`` included) that :func:`write_handoff_buffer` stores under the reference's file name."""
    if handoff_mode not in HANDOFF_MODES:
        raise ValueError(f"handoff_mode must be one of {HANDOFF_MODES}")
    if not torch.cuda.is_available():
        raise _lib.KinError("the handoff-state builder needs a CUDA device; there is no CPU fallback")
    device = torch.device(device)
    cur = approach_config.curriculum_config
    if suite is None:
        if cur.enabled and cur.stages:
            suite, scope = build_curriculum_local_eval_suite(approach_config, seed=seed, stage_index=stage_index, n_episodes=episodes), "curriculum_region"
        else:
            suite, scope = build_fixed_eval_suite(seed=seed, n_episodes=episodes, joint_specs=approach_config.joint_specs,
                                                  start_margin_fraction=approach_config.start_sample_margin_fraction,
                                                  goal_margin_fraction=approach_config.goal_sample_margin_fraction), "fixed_random"
    else:
        scope = "caller_suite"
    n = len(suite)
    stride = (n + 31) // 32 * 32
    t = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=device)  # noqa: E731
    p = lambda x: None if x is None else x.data_ptr()  # noqa: E731
    with torch.cuda.device(device):
        params = ParamsHandle(approach_config)
        iq, idq, ipa, gq, gp = t(suite.initial_q), t(suite.initial_dq), t(suite.initial_prev_action), t(suite.goal_q), t(suite.goal_pose6)
        result = torch.zeros((_D("KIN_RES_ROWS"), stride), dtype=torch.int32, device=device)
        out = torch.zeros((_D("KIN_HO_ROWS"), stride), dtype=torch.float32, device=device)
        steps = torch.zeros(1, dtype=torch.int64, device=device)
        _lib.check(_lib.lib().kin_rollout_handoff_states(params.handle, ctypes.byref(approach_policy.c), p(iq), p(idq), p(ipa), p(gq), p(gp), n, stride,
                                                         int(handoff_confirm_steps), result.data_ptr(), out.data_ptr(), steps.data_ptr(),
                                                         torch.cuda.current_stream().cuda_stream))
        o = out[:, :n].double().cpu().numpy()
    rows = lambda name, w=1: o[_D(name):_D(name) + w].T if w > 1 else o[_D(name)]  # noqa: E731
    fm, sm = rows("KIN_HO_FINAL_METRICS", 4), rows("KIN_HO_SNAP_METRICS", 4)
    final_ready = _finisher_ready(approach_config.reward_config, fm[:, 0], fm[:, 1], fm[:, 2], fm[:, 3])
    snap_step = rows("KIN_HO_SNAP_STEP").astype(int)
    if handoff_mode == "final_settled":
        stored, use_snap = final_ready, False
    elif handoff_mode == "first_confirmed":
        stored, use_snap = snap_step >= 0, True
    else:
        stored, use_snap = np.ones(n, dtype=bool), False
    pre = "KIN_HO_SNAP_" if use_snap else "KIN_HO_FINAL_"
    m = sm if use_snap else fm
    full = HandoffBuffer(episode_id=np.arange(n), step_index=snap_step if use_snap else rows("KIN_HO_FINAL_STEP").astype(int),
                         initial_q=rows(pre + "Q", 7), initial_dq=rows(pre + "DQ", 7), initial_prev_action=rows(pre + "PA", 7),
                         goal_q=rows("KIN_HO_GOAL_Q", 7), goal_pose6=rows("KIN_HO_GOAL_POSE", 6), position_error_norm=m[:, 0],
                         orientation_error_norm=m[:, 1], action_l2=m[:, 2], dq_norm=m[:, 3])
    buf = full.select(stored)
    mean = lambda a: float(np.mean(a)) if len(a) else None  # noqa: E731
    summary = {
        "source_approach_checkpoint": source_checkpoint_name, "approach_algorithm": "ppo", "handoff_mode": handoff_mode, "eval_scope": scope,
        "episode_count": n, "stored_handoff_count": len(buf), "stored_handoff_rate": float(len(buf) / n) if n else 0.0,
        "mean_position_error": mean(buf.position_error_norm), "mean_orientation_error": mean(buf.orientation_error_norm),
        "mean_action_l2": mean(buf.action_l2), "mean_dq_norm": mean(buf.dq_norm), "env_steps": int(steps.item()),
        "states": buf.to_states(dwell_count=int(approach_config.dwell_steps_target), source_checkpoint_name=source_checkpoint_name, handoff_mode=handoff_mode),
        "episode_summaries": [{"episode_id": int(i), "stored_handoff": bool(stored[i]), "final_ready": bool(final_ready[i]),
                               "final_position_error": float(fm[i, 0]), "final_orientation_error": float(fm[i, 1]),
                               "final_action_magnitude": float(fm[i, 2]), "final_dq_norm": float(fm[i, 3])} for i in range(n)],
    }
    return buf, summary


def write_handoff_buffer(artifact_root: str | Path, summary: dict[str, Any]) -> Path:
    """``finisher_handoff_state_buffer.json`` under ``artifact_root`` (the reference's file name, :141)."""
    root = Path(artifact_root)
    root.mkdir(parents=True, exist_ok=True)
    path = root / "finisher_handoff_state_buffer.json"
    path.write_text(json.dumps(summary, indent=2))
    return path
