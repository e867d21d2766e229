"""A stable-baselines3 ``VecEnv`` over the batched GPU env, so the reference's training CLIs run on it unchanged.

The reference builds ``SubprocVecEnv([lambda: ArmKinematicEnv(cfg)] * n)`` and hands it to ``PPO("MultiInputPolicy", env)``
(``kinematic_phase1/train_workspace_expansion.py:180-199``, ``training/train_approach_policy.py:99-114``); its callbacks talk to
the envs through ``training_env.env_method("set_curriculum_stage", i)`` / ``env_method("apply_dock_training_stage", payload)``
(``training/callbacks.py:55,69,151``) and read ``infos[i]["success"]`` on ``dones`` (:73-82).  :class:`KinVecEnv` implements that
protocol -- ``reset / step_async / step_wait / step / close / seed / env_method / get_attr / set_attr / env_is_wrapped`` with
SB3's semantics (auto-reset, ``infos[i]["terminal_observation"]``, ``infos[i]["TimeLimit.truncated"]``) -- on ONE
``BatchedArmKinematicEnv`` (one ``kin_env_step`` launch per step for all envs).  If stable-baselines3 is importable the class
derives from its ``VecEnv`` (and uses gymnasium spaces); otherwise it is a duck-typed stand-in with the same methods, which is
what the tests exercise in this image (SB3 is not installed here).

This adapter is for *compatibility*: SB3's own rollout loop stays on the host (numpy observations, one policy call per step).
The fast path is :class:`rl_brain_trainer_b200.ppo.PPOTrainer` (fused on-device collection and update).
"""

from __future__ import annotations

from typing import Any, Sequence

import numpy as np
import torch

from . import _lib
from .config import Phase1EnvConfig
from .env import OBS_KEYS, OBS_SLICES, BatchedArmKinematicEnv, build_action_space, build_observation_space

try:  # pragma: no cover - stable-baselines3 / gymnasium are absent from the build image
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase
    import gymnasium as _gym

    _HAVE_SB3 = True
except Exception:  # noqa: BLE001
    _VecEnvBase = object
    _gym = None
    _HAVE_SB3 = False

_D = _lib.define
_INFO_SCALARS = ("position_error_norm", "orientation_error_norm", "min_position_error", "action_l2", "executed_delta_q_l2", "delta_q_change_l2")
_REASONS = ("running", "success", "max_steps", "invalid_state")        # termination.py:20-57


def _spaces(n_joints: int):
    if _HAVE_SB3:  # pragma: no cover
        shim = build_observation_space(n_joints)
        obs = _gym.spaces.Dict({k: _gym.spaces.Box(low=float(np.min(b.low)), high=float(np.max(b.high)), shape=b.shape, dtype=np.float32)
                                for k, b in shim.spaces.items()})
        return obs, _gym.spaces.Box(low=-1.0, high=1.0, shape=(n_joints,), dtype=np.float32)
    return build_observation_space(n_joints), build_action_space(n_joints)


class KinVecEnv(_VecEnvBase):
    """``num_envs`` kinematic envs behind SB3's ``VecEnv`` interface (see the module docstring)."""

    def __init__(self, config: Phase1EnvConfig | None = None, num_envs: int = 1, device: str | torch.device = "cuda", *, seed: int = 0,
                 stage_index: int = 0, handoff_states: torch.Tensor | None = None, info_keys: Sequence[str] = _INFO_SCALARS,
                 graph_step: bool = False) -> None:
        self.env = BatchedArmKinematicEnv(config, num_envs, device, auto_reset=True, seed=seed, host_sampler=False, with_aux=True,
                                          graph_step=graph_step)
        self.env.set_curriculum_stage(stage_index)
        if handoff_states is not None:
            self.env.set_handoff_states(handoff_states)
        obs_space, act_space = _spaces(self.env.config.n_joints)
        if _HAVE_SB3:  # pragma: no cover
            super().__init__(int(num_envs), obs_space, act_space)
        else:
            self.num_envs, self.observation_space, self.action_space = int(num_envs), obs_space, act_space
        self.render_mode = None
        self.metadata = {"render_modes": []}
        self._info_keys = tuple(info_keys)
        self._actions: torch.Tensor | None = None
        self._attrs: dict[str, Any] = {}

    # ------------------------------------------------------------------ VecEnv protocol
    def _obs_dict(self, flat: np.ndarray) -> dict[str, np.ndarray]:
        return {k: np.ascontiguousarray(flat[:, OBS_SLICES[k]]) for k in OBS_KEYS}

    def reset(self) -> dict[str, np.ndarray]:
        obs, _ = self.env.reset()
        return self._obs_dict(obs.cpu().numpy())

    def step_async(self, actions: np.ndarray) -> None:
        a = np.asarray(actions, dtype=np.float32)
        if a.shape != (self.num_envs, 7):
            raise ValueError(f"Expected action shape {(self.num_envs, 7)}, got {a.shape}")
        self._actions = torch.as_tensor(a, device=self.env.device)

    def step_wait(self) -> tuple[dict[str, np.ndarray], np.ndarray, np.ndarray, list[dict[str, Any]]]:
        if self._actions is None:
            raise RuntimeError("step_wait() without step_async()")
        obs, reward, terminated, truncated, info = self.env.step(self._actions)
        self._actions = None
        flat = obs.cpu().numpy()
        term, trunc = terminated.cpu().numpy(), truncated.cpu().numpy()
        dones = term | trunc
        host = {k: info[k].cpu().numpy() for k in self._info_keys if k in info}
        success, reason, stage = info["success"].cpu().numpy(), info["reason_code"].cpu().numpy(), info["stage"].cpu().numpy()
        terminal = info["terminal_observation"].cpu().numpy() if dones.any() else None
        infos: list[dict[str, Any]] = []
        for i in range(self.num_envs):
            d: dict[str, Any] = {k: float(v[i]) for k, v in host.items()}
            d["success"] = bool(success[i])
            d["reason"] = _REASONS[int(reason[i])]
            d["curriculum_stage"] = int(stage[i])
            d["TimeLimit.truncated"] = bool(trunc[i] and not term[i])
            if dones[i]:
                d["terminal_observation"] = {k: terminal[i, OBS_SLICES[k]].copy() for k in OBS_KEYS}
            infos.append(d)
        return self._obs_dict(flat), reward.cpu().numpy().astype(np.float32), dones, infos

    def step(self, actions: np.ndarray):
        self.step_async(actions)
        return self.step_wait()

    def close(self) -> None:
        self.env.close()

    def seed(self, seed: int | None = None) -> list[int | None]:
        if seed is not None:
            self.env._seed = int(seed)
        return [None if seed is None else int(seed) + i for i in range(self.num_envs)]

    def _indices(self, indices: Any) -> list[int]:
        if indices is None:
            return list(range(self.num_envs))
        return [int(indices)] if isinstance(indices, (int, np.integer)) else [int(i) for i in indices]

    def env_method(self, method_name: str, *method_args: Any, indices: Any = None, **method_kwargs: Any) -> list[Any]:
        """The calls the reference's callbacks make; they address all envs at once (one shared device config / sampler)."""
        n = len(self._indices(indices))
        if method_name in ("set_curriculum_stage", "get_curriculum_stage", "set_policy_mode", "apply_dock_training_stage", "current_observation"):
            out = getattr(self.env, method_name)(*method_args, **method_kwargs)
            if method_name == "current_observation":
                flat = out.cpu().numpy()
                return [{k: flat[i, OBS_SLICES[k]].copy() for k in OBS_KEYS} for i in self._indices(indices)]
            return [out] * n
        raise AttributeError(f"KinVecEnv.env_method: '{method_name}' is not part of the kinematic env's interface")

    def get_attr(self, attr_name: str, indices: Any = None) -> list[Any]:
        n = len(self._indices(indices))
        if attr_name in self._attrs:
            return [self._attrs[attr_name]] * n
        if attr_name in ("config", "metadata", "render_mode", "action_space", "observation_space"):
            return [getattr(self.env, attr_name, getattr(self, attr_name, None))] * n
        raise AttributeError(attr_name)

    def set_attr(self, attr_name: str, value: Any, indices: Any = None) -> None:
        self._attrs[attr_name] = value

    def env_is_wrapped(self, wrapper_class: Any, indices: Any = None) -> list[bool]:
        return [False] * len(self._indices(indices))

    def get_images(self) -> list[None]:
        return [None] * self.num_envs

    def render(self, mode: str | None = None) -> None:
        return None


def make_vec_env(config: Phase1EnvConfig, n_envs: int, *, seed: int = 0, device: str | torch.device = "cuda", **kwargs: Any) -> KinVecEnv:
    """Drop-in for the reference's ``_make_env`` + ``SubprocVecEnv`` construction (train_workspace_expansion.py:176-183)."""
    return KinVecEnv(config, n_envs, device, seed=seed, **kwargs)
