"""ctypes binding of the CPU fp64 oracle (``oracle/kin_oracle.{h,c}``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product package
``rl_brain_trainer_b200`` never imports this module.

The structs are built by parsing ``kin_oracle.h`` so the C header stays the single source
of truth for the layout.  ``params_from_config`` accepts anything shaped like the reference's
``Phase1EnvConfig`` (``kinematic_phase1/envs/arm_kinematic_env.py:32-66``) -- the reference's own
dataclass when it is importable, or the product's mirror of it -- by attribute name only.
"""

from __future__ import annotations

import ctypes
import os
import re
import subprocess
from pathlib import Path
from typing import Any, Sequence

import numpy as np

_HERE = Path(__file__).resolve().parent
_HEADER = _HERE / "kin_oracle.h"
_LIB_PATH = _HERE / "_build" / "libkin_oracle.so"

N_JOINTS = 7
OBS_DIM = 56
ROUTE_OBS_DIM = 80
MODE_APPROACH = 0
MODE_DOCK = 1
REASONS = ("running", "success", "max_steps", "invalid_state")

APPROACH_COMPONENT_NAMES = (
    "position_progress", "global_orientation_progress", "near_field_orientation_progress", "orientation_progress",
    "orientation_milestone_bonus", "near_field_orientation_center", "pre_near_goal_bonus", "near_goal_bonus",
    "pre_near_to_near_progress", "near_goal_bonus_scale", "coarse_orientation_bonus", "handover_bonus",
    "handover_retention_bonus", "handover_dwell_bonus", "handover_leave_penalty", "handover_regression_penalty",
    "dock_coarse_ready_bonus", "dock_coarse_ready_retention_bonus", "dock_coarse_ready_dwell_bonus",
    "dock_coarse_ready_leave_penalty", "dock_coarse_ready_regression_penalty", "finisher_ready_bonus",
    "finisher_ready_retention_bonus", "finisher_ready_dwell_bonus", "finisher_ready_leave_penalty",
    "finisher_ready_regression_penalty", "near_handoff_action_penalty", "near_handoff_dq_penalty",
    "near_handoff_motion_bonus", "near_handoff_settle_bonus", "same_step_alignment_bonus", "dwell_bonus",
    "drift_penalty", "near_goal_leave_penalty", "drift_penalty_scale", "near_goal_entry_count",
    "near_goal_drift_count", "smoothness_penalty", "smoothness_multiplier", "joint_limit_penalty", "success_bonus",
    "curr_pos_error", "curr_ori_error", "curr_action_norm", "curr_dq_norm", "dwell_count", "in_pre_near_goal",
    "in_near_goal", "in_handover_zone", "in_dock_coarse_ready", "in_dock_coarse_ready_pose", "in_finisher_ready",
    "in_finisher_ready_pose", "in_near_handoff_zone",
)
DOCK_COMPONENT_NAMES = (
    "position_progress", "orientation_progress", "stay_in_zone_bonus", "dwell_bonus", "working_range_bonus",
    "working_range_dwell_bonus", "tight_pose_bonus", "tight_pose_dwell_bonus", "strict_pose_leave_penalty",
    "strict_center_reward", "strict_center_position_penalty", "strict_center_orientation_penalty",
    "strict_center_small_action_bonus", "strict_center_dwell_bonus", "tight_position_shaping",
    "tight_orientation_shaping", "convergence_position_progress", "convergence_orientation_progress",
    "orientation_position_gate_scale", "entry_action_penalty_scale", "leave_zone_penalty",
    "working_range_exit_penalty", "drift_penalty", "smoothness_penalty", "action_delta_violation_penalty",
    "delta_q_change_penalty", "preserve_state_bonus", "strict_hold_bonus", "low_motion_bonus",
    "tiny_correction_bonus", "worse_than_entry_penalty", "near_strict_regression_penalty",
    "aggressive_action_penalty", "dq_penalty", "joint_limit_penalty", "success_bonus", "basin_outer_bonus",
    "basin_inner_bonus", "basin_dwell_bonus", "basin_outer_exit_penalty", "basin_inner_exit_penalty",
    "basin_dwell_break_penalty", "basin_drift_penalty", "basin_zone_index", "curr_pos_error", "curr_ori_error",
    "dwell_count", "in_tight_pose", "in_near_strict", "entry_pos_error", "entry_ori_error", "entry_action_l2",
    "entry_dq_norm", "entry_to_curr_delta_position_error", "entry_to_curr_delta_orientation_error",
    "entry_to_curr_delta_action_l2", "entry_to_curr_delta_dq_norm", "near_goal_entry_count",
    "near_goal_drift_count", "in_near_goal",
)
ROUTE_COMPONENT_NAMES = (
    "q_goal_progress", "ee_position_progress", "ee_orientation_progress", "route_tangent_progress_bonus",
    "same_step_route_ready_bonus", "route_ready_dwell_bonus", "low_motion_near_waypoint_bonus",
    "orientation_regression_penalty", "q_route_regression_penalty", "off_route_penalty",
    "action_smoothness_penalty", "dq_penalty", "no_progress_penalty", "curr_q_error", "curr_pos_error",
    "curr_ori_error", "route_ready",
)

# --------------------------------------------------------------------------------------
# header -> ctypes
# --------------------------------------------------------------------------------------
_SCALARS = {"double": ctypes.c_double, "int": ctypes.c_int, "float": ctypes.c_float}
_struct_cache: dict[str, type] = {}


def _parse_struct(text: str, name: str) -> type:
    if name in _struct_cache:
        return _struct_cache[name]
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, re.S)
    if m is None:
        raise RuntimeError(f"struct {name} not found in {_HEADER}")
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    fields: list[tuple[str, Any]] = []
    for raw in body.split(";"):
        line = raw.strip()
        if not line:
            continue
        mm = re.match(r"(const\s+)?(\w+)\s*(\*)?\s*(\w+)(?:\[(\d+)\])?$", line)
        if mm is None:
            raise RuntimeError(f"cannot parse field {line!r} of {name}")
        _, ctype, ptr, fname, arr = mm.groups()
        if ptr:
            t: Any = ctypes.c_void_p
        elif ctype in _SCALARS:
            t = _SCALARS[ctype]
        else:
            t = _parse_struct(text, ctype)
        if arr:
            t = t * int(arr)
        fields.append((fname, t))
    cls = type(name, (ctypes.Structure,), {"_fields_": fields})
    _struct_cache[name] = cls
    return cls


_header_text = _HEADER.read_text()
Params = _parse_struct(_header_text, "kor_params")
State = _parse_struct(_header_text, "kor_state")
StepOut = _parse_struct(_header_text, "kor_step_out")
Mlp = _parse_struct(_header_text, "kor_mlp")
EpisodeResult = _parse_struct(_header_text, "kor_episode_result")
Route = _parse_struct(_header_text, "kor_route")
RouteState = _parse_struct(_header_text, "kor_route_state")
RouteStepOut = _parse_struct(_header_text, "kor_route_step_out")

_lib: ctypes.CDLL | None = None


def build(force: bool = False) -> Path:
    """Compile the oracle with gcc (``oracle/Makefile``)."""
    srcs = [_HERE / "kin_oracle.c", _HERE / "kin_oracle_route.c", _HEADER]
    stale = (not _LIB_PATH.exists()) or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        subprocess.run(["make", "-s", "-C", str(_HERE)] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_LIB_PATH))
        dp = ctypes.POINTER(ctypes.c_double)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int)
        L.kor_fk_pose6.argtypes = [dp, dp]
        L.kor_fk_matrix.argtypes = [dp, dp]
        L.kor_wrap_to_pi.argtypes = [ctypes.c_double]
        L.kor_wrap_to_pi.restype = ctypes.c_double
        L.kor_default_joint_specs.argtypes = [dp, dp, dp]
        L.kor_reset.argtypes = [ctypes.POINTER(Params), ctypes.POINTER(State), ctypes.c_int, dp, dp, dp, dp, dp]
        L.kor_observation.argtypes = [ctypes.POINTER(Params), ctypes.POINTER(State), fp]
        L.kor_step.argtypes = [ctypes.POINTER(Params), ctypes.POINTER(State), dp, ctypes.POINTER(StepOut), fp]
        L.kor_step_batch.argtypes = [ctypes.POINTER(Params), ctypes.POINTER(State), dp, ctypes.c_int,
                                     ctypes.POINTER(StepOut), fp]
        L.kor_mlp_forward.argtypes = [ctypes.POINTER(Mlp), fp, fp, fp]
        L.kor_eval_approach_finisher.argtypes = [ctypes.POINTER(Params), ctypes.POINTER(Params), ctypes.POINTER(Mlp),
                                                 ctypes.POINTER(Mlp), dp, dp, dp, dp, dp, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, ctypes.POINTER(EpisodeResult),
                                                 ctypes.POINTER(ctypes.c_longlong)]
        L.kor_route_build.argtypes = [dp, ctypes.c_int, dp, dp, dp]
        L.kor_route_reset.argtypes = [ctypes.POINTER(Params), ctypes.POINTER(Route), ctypes.POINTER(RouteState),
                                      ctypes.c_int, ctypes.c_int, dp, dp, dp]
        L.kor_route_observation.argtypes = [ctypes.POINTER(Params), ctypes.POINTER(Route), ctypes.POINTER(RouteState), fp]
        L.kor_route_step.argtypes = [ctypes.POINTER(Params), ctypes.POINTER(Route), ctypes.POINTER(RouteState), dp,
                                     ctypes.c_int, ctypes.c_int, ctypes.POINTER(RouteStepOut), fp]
        L.kor_route_sequential_probe.argtypes = [ctypes.POINTER(Params), ctypes.POINTER(Route), ctypes.POINTER(Mlp),
                                                 ctypes.c_int, ctypes.c_int, dp, ip, dp,
                                                 ctypes.POINTER(ctypes.c_longlong)]
        L.kor_route_sequential_probe.restype = ctypes.c_int
        _lib = L
    return _lib


def _dptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _fptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _f64(x: Any, shape: tuple[int, ...] | None = None) -> np.ndarray | None:
    if x is None:
        return None
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if shape is not None:
        a = a.reshape(shape)
    return a


# --------------------------------------------------------------------------------------
# config flattening
# --------------------------------------------------------------------------------------
def _get(obj: Any, name: str, default: Any = None) -> Any:
    if isinstance(obj, dict):
        return obj.get(name, default)
    return getattr(obj, name, default)


def params_from_config(cfg: Any, route_reward_config: Any | None = None) -> Any:
    """Flatten a ``Phase1EnvConfig``-shaped object into ``kor_params`` (by attribute name)."""
    p = Params()
    specs = _get(cfg, "joint_specs")
    for i, spec in enumerate(specs):
        p.joint_lower[i] = float(_get(spec, "lower"))
        p.joint_upper[i] = float(_get(spec, "upper"))
        p.joint_delta_limit[i] = float(_get(spec, "delta_limit"))
    names = {f[0] for f in Params._fields_}
    for fname, ftype in Params._fields_:
        if fname.startswith(("joint_", "ar_", "dr_", "rr_", "term_", "obs_")):
            continue
        v = _get(cfg, fname)
        if v is None:
            raise KeyError(f"config lacks {fname}")
        setattr(p, fname, int(v) if ftype is ctypes.c_int else float(v))
    term = _get(cfg, "termination_config")
    p.term_max_episode_steps = int(_get(term, "max_episode_steps"))
    p.term_success_pos_threshold_m = float(_get(term, "success_pos_threshold_m"))
    p.term_success_ori_threshold_rad = float(_get(term, "success_ori_threshold_rad"))
    p.term_success_dwell_steps = int(_get(term, "success_dwell_steps"))
    p.term_require_orientation = int(bool(_get(term, "require_orientation")))
    p.term_terminate_on_success = int(bool(_get(term, "terminate_on_success")))
    oc = _get(cfg, "observation_config")
    p.obs_pos_err_scale_m = float(_get(oc, "pos_err_scale_m"))
    p.obs_ori_err_scale_rad = float(_get(oc, "ori_err_scale_rad"))

    def fill(prefix: str, sub: Any, skip: Sequence[str] = ()) -> None:
        for fname, ftype in Params._fields_:
            if not fname.startswith(prefix) or fname in skip:
                continue
            v = _get(sub, fname[len(prefix):])
            if v is None:
                raise KeyError(f"config lacks {fname}")
            setattr(p, fname, int(v) if ftype is ctypes.c_int else float(v))

    ar = _get(cfg, "reward_config")
    fill("ar_", ar, skip=("ar_n_milestones", "ar_orientation_milestone_thresholds_rad", "ar_orientation_milestone_bonuses"))
    thr = tuple(_get(ar, "orientation_milestone_thresholds_rad", ()) or ())
    bon = tuple(_get(ar, "orientation_milestone_bonuses", ()) or ())
    n = min(len(thr), len(bon))  # zip(strict=False), reward_approach.py:111
    if n > 4:
        raise ValueError("oracle supports at most 4 orientation milestones")
    p.ar_n_milestones = n
    for i in range(n):
        p.ar_orientation_milestone_thresholds_rad[i] = float(thr[i])
        p.ar_orientation_milestone_bonuses[i] = float(bon[i])
    fill("dr_", _get(cfg, "dock_reward_config"))
    if route_reward_config is not None:
        fill("rr_", route_reward_config)
    assert names  # silence linters
    return p


def default_joint_specs() -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    lo, hi, dl = (np.zeros(7) for _ in range(3))
    lib().kor_default_joint_specs(_dptr(lo), _dptr(hi), _dptr(dl))
    return lo, hi, dl


# --------------------------------------------------------------------------------------
# kinematics
# --------------------------------------------------------------------------------------
def fk_pose6(q: Any) -> np.ndarray:
    q = _f64(q)
    single = q.ndim == 1
    q2 = np.ascontiguousarray(q.reshape(-1, 7))
    out = np.zeros((q2.shape[0], 6))
    L = lib()
    for i in range(q2.shape[0]):
        L.kor_fk_pose6(_dptr(q2[i]), _dptr(out[i]))
    return out[0] if single else out


def fk_matrix(q: Any) -> np.ndarray:
    q = _f64(q, (7,))
    T = np.zeros(16)
    lib().kor_fk_matrix(_dptr(q), _dptr(T))
    return T.reshape(4, 4)


def wrap_to_pi(v: float) -> float:
    return float(lib().kor_wrap_to_pi(float(v)))


# --------------------------------------------------------------------------------------
# single env (mirrors ArmKinematicEnv for the explicit-options path)
# --------------------------------------------------------------------------------------
class OracleEnv:
    def __init__(self, params: Any) -> None:
        self.params = params
        self.state = State()
        self._L = lib()

    def reset(self, *, mode: int, initial_q, goal_q=None, goal_pose6=None, initial_dq=None, initial_prev_action=None):
        iq, gq, gp = _f64(initial_q, (7,)), _f64(goal_q, (7,)), _f64(goal_pose6, (6,))
        idq, ipa = _f64(initial_dq, (7,)), _f64(initial_prev_action, (7,))
        self._L.kor_reset(ctypes.byref(self.params), ctypes.byref(self.state), int(mode), _dptr(iq), _dptr(idq),
                          _dptr(ipa), _dptr(gq), _dptr(gp))
        return self.observation()

    def observation(self) -> np.ndarray:
        obs = np.zeros(56, dtype=np.float32)
        self._L.kor_observation(ctypes.byref(self.params), ctypes.byref(self.state), _fptr(obs))
        return obs

    def step(self, action):
        a = _f64(action, (7,))
        out = StepOut()
        obs = np.zeros(56, dtype=np.float32)
        self._L.kor_step(ctypes.byref(self.params), ctypes.byref(self.state), _dptr(a), ctypes.byref(out), _fptr(obs))
        return obs, out


def state_array(n: int):
    return (State * n)()


def step_batch(params, states, actions: np.ndarray):
    n = len(states)
    a = _f64(actions, (n, 7))
    outs = (StepOut * n)()
    obs = np.zeros((n, 56), dtype=np.float32)
    lib().kor_step_batch(ctypes.byref(params), states, _dptr(a), n, outs, _fptr(obs))
    return obs, outs


# --------------------------------------------------------------------------------------
# policy
# --------------------------------------------------------------------------------------
class OracleMlp:
    """Holds fp32 copies of an SB3 ``policy.pth`` state dict (key names per SURVEY F4)."""

    KEYS = {
        "pi_w0": "mlp_extractor.policy_net.0.weight", "pi_b0": "mlp_extractor.policy_net.0.bias",
        "pi_w1": "mlp_extractor.policy_net.2.weight", "pi_b1": "mlp_extractor.policy_net.2.bias",
        "act_w": "action_net.weight", "act_b": "action_net.bias",
        "vf_w0": "mlp_extractor.value_net.0.weight", "vf_b0": "mlp_extractor.value_net.0.bias",
        "vf_w1": "mlp_extractor.value_net.2.weight", "vf_b1": "mlp_extractor.value_net.2.bias",
        "val_w": "value_net.weight", "val_b": "value_net.bias",
    }

    def __init__(self, weights: dict[str, np.ndarray]) -> None:
        self._keep: dict[str, np.ndarray] = {}
        self.c = Mlp()
        for field, key in self.KEYS.items():
            if key in weights:
                arr = np.ascontiguousarray(np.asarray(weights[key], dtype=np.float32))
                self._keep[field] = arr
                setattr(self.c, field, arr.ctypes.data)
        self.c.in_dim = int(self._keep["pi_w0"].shape[1])
        self.c.has_value = int("vf_w0" in self._keep)

    def forward(self, obs: np.ndarray) -> tuple[np.ndarray, float]:
        o = np.ascontiguousarray(obs, dtype=np.float32)
        a = np.zeros(7, dtype=np.float32)
        v = np.zeros(1, dtype=np.float32)
        lib().kor_mlp_forward(ctypes.byref(self.c), _fptr(o), _fptr(a), _fptr(v))
        return a, float(v[0])


# --------------------------------------------------------------------------------------
# Approach -> Finisher evaluation
# --------------------------------------------------------------------------------------
_RESULT_FIELDS = [f[0] for f in EpisodeResult._fields_ if f[0] not in ("pad0",)]


def eval_approach_finisher(pa, pf, approach: OracleMlp, finisher: OracleMlp | None, *, initial_q, goal_q,
                           goal_pose6=None, initial_dq=None, initial_prev_action=None, handoff_confirm_steps: int = 2,
                           n_threads: int = 0) -> tuple[dict[str, np.ndarray], int]:
    iq = _f64(initial_q)
    n = iq.shape[0]
    gq, gp = _f64(goal_q, (n, 7)), _f64(goal_pose6, (n, 6)) if goal_pose6 is not None else None
    idq = _f64(initial_dq, (n, 7)) if initial_dq is not None else None
    ipa = _f64(initial_prev_action, (n, 7)) if initial_prev_action is not None else None
    res = (EpisodeResult * n)()
    steps = ctypes.c_longlong(0)
    lib().kor_eval_approach_finisher(ctypes.byref(pa), ctypes.byref(pf) if pf is not None else None,
                                     ctypes.byref(approach.c), ctypes.byref(finisher.c) if finisher is not None else None,
                                     _dptr(iq), _dptr(idq), _dptr(ipa), _dptr(gq), _dptr(gp), n,
                                     int(handoff_confirm_steps), int(n_threads), res, ctypes.byref(steps))
    raw = np.frombuffer(res, dtype=np.dtype(EpisodeResult))   # structured view of the C array (65 536 episodes convert in milliseconds)
    out: dict[str, np.ndarray] = {name: np.array(raw[name]) for name in _RESULT_FIELDS}
    return out, int(steps.value)


# --------------------------------------------------------------------------------------
# route
# --------------------------------------------------------------------------------------
class OracleRoute:
    def __init__(self, q_goals: np.ndarray) -> None:
        self.q_goal = _f64(q_goals)
        n = self.q_goal.shape[0]
        self.pose6 = np.zeros((n, 6))
        self.next_q_delta = np.zeros((n, 7))
        self.progress_m = np.zeros(n)
        lib().kor_route_build(_dptr(self.q_goal), n, _dptr(self.pose6), _dptr(self.next_q_delta), _dptr(self.progress_m))
        self.c = Route()
        self.c.n_waypoints = n
        self.c.q_goal = self.q_goal.ctypes.data
        self.c.pose6 = self.pose6.ctypes.data
        self.c.next_q_delta = self.next_q_delta.ctypes.data
        self.c.progress_m = self.progress_m.ctypes.data

    def __len__(self) -> int:
        return int(self.c.n_waypoints)


class OracleRouteEnv:
    def __init__(self, params, route: OracleRoute, *, sequence_length: int = 0, reset_ready_streak_on_advance: bool = True,
                 max_route_index: int | None = None) -> None:
        self.params, self.route = params, route
        self.sequence_length = int(sequence_length)
        self.reset_streak = bool(reset_ready_streak_on_advance)
        self.max_route_index = len(route) - 1 if max_route_index is None else int(max_route_index)
        self.state = RouteState()
        self._L = lib()

    def reset(self, *, route_index: int, start_route_index: int | None = None, initial_q=None, initial_dq=None,
              initial_prev_action=None) -> np.ndarray:
        start = max(route_index - 1, 0) if start_route_index is None else int(start_route_index)
        ri = int(route_index)
        if self.sequence_length > 0:
            max_index = min(self.max_route_index, len(self.route) - 1)
            ri = int(np.clip(ri, 1, max_index))
        self._L.kor_route_reset(ctypes.byref(self.params), ctypes.byref(self.route.c), ctypes.byref(self.state), ri, start,
                                _dptr(_f64(initial_q, (7,))), _dptr(_f64(initial_dq, (7,))),
                                _dptr(_f64(initial_prev_action, (7,))))
        if self.sequence_length > 0:
            self.state.last_route_index = int(min(max_index, ri + max(self.sequence_length, 1) - 1))
        return self.observation()

    def observation(self) -> np.ndarray:
        obs = np.zeros(80, dtype=np.float32)
        self._L.kor_route_observation(ctypes.byref(self.params), ctypes.byref(self.route.c), ctypes.byref(self.state), _fptr(obs))
        return obs

    def step(self, action):
        a = _f64(action, (7,))
        out = RouteStepOut()
        obs = np.zeros(80, dtype=np.float32)
        self._L.kor_route_step(ctypes.byref(self.params), ctypes.byref(self.route.c), ctypes.byref(self.state), _dptr(a),
                               int(self.sequence_length > 0), int(self.reset_streak), ctypes.byref(out), _fptr(obs))
        return obs, out


def route_sequential_probe(params, route: OracleRoute, policy: OracleMlp, *, start_index: int = 1,
                           end_index: int | None = None, start_q=None):
    end = len(route) - 1 if end_index is None else int(end_index)
    m = end - start_index + 1
    flags = np.zeros(m, dtype=np.int32)
    errs = np.zeros((m, 3))
    steps = ctypes.c_longlong(0)
    prefix = lib().kor_route_sequential_probe(ctypes.byref(params), ctypes.byref(route.c), ctypes.byref(policy.c),
                                              int(start_index), end, _dptr(_f64(start_q, (7,))),
                                              flags.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _dptr(errs),
                                              ctypes.byref(steps))
    return int(prefix), flags, errs, int(steps.value)


__all__ = [name for name in dir() if not name.startswith("_")]
if os.environ.get("KIN_ORACLE_EAGER_BUILD"):
    build()
