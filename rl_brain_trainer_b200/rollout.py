"""Fused policy-in-loop Approach -> Finisher evaluation (one kernel launch per suite) and its host-side summaries.

``evaluate_suite`` replaces the per-episode Python loop of ``evaluate_workspace_expansion_checkpoint``
(``kinematic_phase1/eval/eval_workspace_expansion.py:86-211``) and ``_run_pairs``
(``eval/eval_full_workspace_coverage.py:120-190``); the summary helpers reproduce ``_failure_reason`` /
``_summarize_stage`` (``eval_workspace_expansion.py:47-83``) and the gate score
(``workspace/workspace_curriculum.py:58-90``) on device-reduced statistics.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from . import _lib
from .config import Phase1EnvConfig
from .env import ParamsHandle
from .policy import PolicyWeights
from .samplers import EvalSuite

VARIANT_FFMA = 0   # strict fp32 on the FP32 pipe
VARIANT_TC = 1     # MLP on the 5th-gen tensor cores (tcgen05 kind::f16: fp16 operands, fp32 accumulators in TMEM), csrc/kin_rollout_tc16.cu


def _D(name: str) -> int:
    return _lib.define(name)


@dataclass
class RolloutResult:
    """Device-resident per-episode rows (see KIN_RES_* in include/kin_b200.h); properties are views."""

    raw: torch.Tensor          # [KIN_RES_ROWS, stride] int32 words
    n: int
    env_steps: torch.Tensor    # [1] int64 on device

    def _u(self, row: str) -> torch.Tensor:
        return self.raw[_D(row), : self.n]

    def _f(self, row: str) -> torch.Tensor:
        return self.raw[_D(row), : self.n].view(torch.float32)

    @property
    def success(self) -> torch.Tensor:
        return self._u("KIN_RES_SUCCESS") != 0

    @property
    def flags(self) -> torch.Tensor:
        return self._u("KIN_RES_FLAGS")

    approach_success = property(lambda self: (self.flags & 1) != 0)
    ready_hit = property(lambda self: (self.flags & 2) != 0)
    ready_dwell = property(lambda self: (self.flags & 4) != 0)
    final_ready = property(lambda self: (self.flags & 8) != 0)
    handoff_kind = property(lambda self: (self.flags >> 4) & 3)
    handoff_step = property(lambda self: self._u("KIN_RES_HANDOFF_STEP"))
    first_ready_step = property(lambda self: self._u("KIN_RES_FIRST_READY_STEP"))
    max_ready_streak = property(lambda self: self._u("KIN_RES_MAX_READY_STREAK"))
    approach_steps = property(lambda self: self._u("KIN_RES_STEPS") & 0xFFFF)
    finisher_steps = property(lambda self: (self._u("KIN_RES_STEPS") >> 16) & 0xFFFF)
    final_position_error = property(lambda self: self._f("KIN_RES_FINAL_POS"))
    final_orientation_error = property(lambda self: self._f("KIN_RES_FINAL_ORI"))
    approach_final_position_error = property(lambda self: self._f("KIN_RES_APPROACH_POS"))
    approach_final_orientation_error = property(lambda self: self._f("KIN_RES_APPROACH_ORI"))
    min_position_error = property(lambda self: self._f("KIN_RES_MIN_POS"))
    min_orientation_error = property(lambda self: self._f("KIN_RES_MIN_ORI"))
    approach_final_action_magnitude = property(lambda self: self._f("KIN_RES_APPROACH_ACTION"))
    approach_final_dq_norm = property(lambda self: self._f("KIN_RES_APPROACH_DQ"))
    final_action_magnitude = property(lambda self: self._f("KIN_RES_FINAL_ACTION"))
    final_dq_norm = property(lambda self: self._f("KIN_RES_FINAL_DQ"))

    @property
    def final_q(self) -> torch.Tensor:
        r = _D("KIN_RES_FINAL_Q")
        return self.raw[r:r + 7, : self.n].view(torch.float32).t()

    def to_numpy(self) -> dict[str, np.ndarray]:
        names = ("success", "approach_success", "ready_hit", "ready_dwell", "final_ready", "handoff_kind", "handoff_step",
                 "first_ready_step", "max_ready_streak", "approach_steps", "finisher_steps", "final_position_error",
                 "final_orientation_error", "approach_final_position_error", "approach_final_orientation_error",
                 "min_position_error", "min_orientation_error", "final_action_magnitude", "final_dq_norm", "final_q")
        return {k: getattr(self, k).detach().cpu().numpy() for k in names}


class ApproachFinisherRollout:
    """Holds the two env parameter handles + policies and launches the fused rollout kernel."""

    def __init__(self, approach_config: Phase1EnvConfig, approach_policy: PolicyWeights, finisher_config: Phase1EnvConfig | None = None,
                 finisher_policy: PolicyWeights | None = None, *, device: str | torch.device = "cuda", handoff_confirm_steps: int = 2,
                 variant: int = VARIANT_FFMA) -> None:
        if not torch.cuda.is_available():
            raise _lib.KinError("the fused rollout needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            self.pa = ParamsHandle(approach_config)
            self.pf = ParamsHandle(finisher_config) if (finisher_config is not None and finisher_policy is not None) else None
        self.approach_policy, self.finisher_policy = approach_policy, finisher_policy
        self.confirm = int(handoff_confirm_steps)
        self.variant = int(variant)
        self.max_steps_per_episode = approach_config.termination_config.max_episode_steps + (
            finisher_config.termination_config.max_episode_steps if self.pf is not None else 0)

    def upload(self, suite: EvalSuite, *, pinned: bool = False) -> dict[str, torch.Tensor | None]:
        """Host suite -> device tensors (float32).  ``pinned=True`` stages through page-locked memory (bench e2e leg)."""
        out: dict[str, torch.Tensor | None] = {}
        for name in ("initial_q", "initial_dq", "initial_prev_action", "goal_q", "goal_pose6"):
            arr = getattr(suite, name)
            if arr is None:
                out[name] = None
                continue
            t = torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float32))
            if pinned:
                t = t.pin_memory()
            out[name] = t.to(self.device, non_blocking=pinned)
        return out

    def run(self, dev: dict[str, torch.Tensor | None], *, out: RolloutResult | None = None) -> RolloutResult:
        iq = dev["initial_q"]
        n = int(iq.shape[0])
        stride = (n + 31) // 32 * 32
        if out is None:
            out = RolloutResult(raw=torch.zeros((_D("KIN_RES_ROWS"), stride), dtype=torch.int32, device=self.device), n=n,
                                env_steps=torch.zeros(1, dtype=torch.int64, device=self.device))
        p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().kin_rollout_approach_finisher(
                self.pa.handle, None if self.pf is None else self.pf.handle, ctypes.byref(self.approach_policy.c),
                None if self.finisher_policy is None or self.pf is None else ctypes.byref(self.finisher_policy.c),
                p(iq), p(dev.get("initial_dq")), p(dev.get("initial_prev_action")), p(dev.get("goal_q")), p(dev.get("goal_pose6")),
                n, stride, self.confirm, self.variant, out.raw.data_ptr(), out.env_steps.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return out

    def evaluate_suite(self, suite: EvalSuite) -> RolloutResult:
        return self.run(self.upload(suite))

    def evaluate_stream(self, host_inputs: list[dict[str, torch.Tensor | None]], host_results: list[torch.Tensor]) -> list[torch.Tensor]:
        """Evaluate a sequence of suites held in PINNED host memory, results into pinned host tensors, copies overlapped.

        Three streams, two device slots: the H2D copy of suite k+1 and the D2H copy of result k-1 run while the rollout
        kernel of suite k computes.  ``host_inputs[k]`` maps ``initial_q / goal_q / ...`` to pinned ``[n,7|6]`` float32
        tensors (or None); ``host_results[k]`` is a pinned int32 ``[KIN_RES_ROWS, stride]`` tensor.  Returns per-suite
        device ``env_steps`` counters; the caller synchronises (e.g. ``torch.cuda.synchronize``) before reading results.
        """
        dev = self.device
        if not hasattr(self, "_streams"):
            self._streams = [torch.cuda.Stream(dev) for _ in range(3)]
        s_in, s_run, s_out = self._streams
        cur = torch.cuda.current_stream(dev)
        for st in self._streams:
            st.wait_stream(cur)
        n = int(host_inputs[0]["initial_q"].shape[0])
        stride = (n + 31) // 32 * 32
        slots = []
        for _ in range(2):
            d_in = {k: (None if v is None else torch.empty(v.shape, dtype=torch.float32, device=dev)) for k, v in host_inputs[0].items()}
            res = RolloutResult(raw=torch.zeros((_D("KIN_RES_ROWS"), stride), dtype=torch.int32, device=dev), n=n,
                                env_steps=torch.zeros(1, dtype=torch.int64, device=dev))
            slots.append({"in": d_in, "res": res, "free": None})
        counters = []
        for k, (h_in, h_out) in enumerate(zip(host_inputs, host_results)):
            slot = slots[k % 2]
            with torch.cuda.stream(s_in):
                if slot["free"] is not None:
                    s_in.wait_event(slot["free"])            # the previous user of this slot has been copied out
                for key, v in h_in.items():
                    if v is not None:
                        slot["in"][key].copy_(v, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            with torch.cuda.stream(s_run):
                s_run.wait_event(ev_in)
                slot["res"].env_steps.zero_()
                self.run(slot["in"], out=slot["res"])
                ev_run = torch.cuda.Event()
                ev_run.record(s_run)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_run)
                h_out.copy_(slot["res"].raw, non_blocking=True)
                steps = slot["res"].env_steps.clone()
                ev_out = torch.cuda.Event()
                ev_out.record(s_out)
            slot["free"] = ev_out
            counters.append(steps)
        for st in self._streams:
            cur.wait_stream(st)
        return counters


# ----------------------------------------------------------------------------------------------
# summaries
# ----------------------------------------------------------------------------------------------
FAILURE_REASONS = ("success", "position", "orientation", "motion_action", "motion_dq", "dwell", "timeout_or_regression")


def failure_reason_codes(result: RolloutResult, approach_config: Phase1EnvConfig, handoff_confirm_steps: int = 2) -> torch.Tensor:
    """``_failure_reason`` / ``_reason`` (eval_workspace_expansion.py:47-66, eval_full_workspace_coverage.py:38-52) per episode, as an
    index into ``FAILURE_REASONS``: the checks read the APPROACH result in the reference's order (position, orientation, action, dq,
    dock-coarse-ready dwell)."""
    rc = approach_config.reward_config
    reason = torch.full((result.n,), 6, dtype=torch.int64, device=result.raw.device)
    reason = torch.where(result.max_ready_streak < int(handoff_confirm_steps), torch.full_like(reason, 5), reason)
    reason = torch.where(result.approach_final_dq_norm > rc.finisher_ready_dq_threshold, torch.full_like(reason, 4), reason)
    reason = torch.where(result.approach_final_action_magnitude > rc.finisher_ready_action_threshold, torch.full_like(reason, 3), reason)
    reason = torch.where(result.approach_final_orientation_error > rc.finisher_ready_ori_threshold_rad, torch.full_like(reason, 2), reason)
    reason = torch.where(result.approach_final_position_error > rc.finisher_ready_pos_threshold_m, torch.full_like(reason, 1), reason)
    return torch.where(result.success, torch.zeros_like(reason), reason)


def summarize(result: RolloutResult, approach_config: Phase1EnvConfig, handoff_confirm_steps: int = 2) -> dict[str, Any]:
    """``_summarize_stage`` + ``_failure_reason`` (eval_workspace_expansion.py:47-83) from device reductions."""
    ok = result.success
    reason = failure_reason_codes(result, approach_config, handoff_confirm_steps)
    counts = torch.bincount(reason, minlength=len(FAILURE_REASONS)).cpu().numpy()
    regress = (result.approach_final_position_error > result.min_position_error + 0.002) | \
              (result.approach_final_orientation_error > result.min_orientation_error + 0.01)
    f = lambda t: float(t.float().mean().item())  # noqa: E731
    return {
        "episode_count": result.n, "success_rate": f(ok), "finisher_ready_hit_rate": f(result.ready_hit),
        "dwell_success_rate": f(result.ready_dwell), "mean_final_position_error": f(result.final_position_error),
        "mean_final_orientation_error": f(result.final_orientation_error), "mean_final_action_magnitude": f(result.final_action_magnitude),
        "mean_final_dq_norm": f(result.final_dq_norm), "regression_rate": f(regress),
        "failure_reason_counts": {FAILURE_REASONS[i]: int(c) for i, c in enumerate(counts) if c},
        "env_steps": int(result.env_steps.item()),
    }
