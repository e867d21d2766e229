"""Shared helpers for the test-suite (oracle construction from presets, fixtures)."""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from oracle import kin_oracle as ko
from rl_brain_trainer_b200 import config as kcfg

GOLD = Path(__file__).resolve().parent / "golden"
POLICY_DIR = kcfg.PRESET_DIR / "policies"


def golden(name: str):
    return np.load(GOLD / name, allow_pickle=False)


def env_config(preset: str) -> kcfg.Phase1EnvConfig:
    return kcfg.load_preset(preset)


def env_config_from_json(path: Path) -> kcfg.Phase1EnvConfig:
    return kcfg.to_env_config(json.loads(Path(path).read_text()))


def oracle_params(cfg, route_reward=None):
    return ko.params_from_config(cfg, route_reward)


def policy_weights(name: str) -> dict[str, np.ndarray]:
    with np.load(POLICY_DIR / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


def oracle_policy(name: str) -> ko.OracleMlp:
    return ko.OracleMlp(policy_weights(name))


from oracle.parity import scale_decision_thresholds, threshold_sensitive_episodes  # noqa: E402,F401  (re-exported for the tests)
