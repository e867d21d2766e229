"""Differential test of the C oracle against the LIVE reference (only where /root/reference exists, i.e. the build container)."""

from __future__ import annotations

import sys

import numpy as np
import pytest

from oracle import kin_oracle as ko

from .conftest import REFERENCE_PKG

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    sys.dont_write_bytecode = True
    if str(REFERENCE_PKG) not in sys.path:
        sys.path.insert(0, str(REFERENCE_PKG))
    import hrl_trainer.kinematic_phase1.envs.arm_kinematic_env as m

    return m


def test_random_open_loop_against_live_reference(ref):
    from rl_brain_trainer_b200 import config as kcfg
    from hrl_trainer.kinematic_phase1.training.policy_config import to_env_config

    rng = np.random.default_rng(2026)
    for preset, mode, mode_idx in (("approach_dynamic_scale_big", "approach", 0), ("finisher_noop_ft", "dock", 1)):
        d = kcfg.preset_dict(preset)
        ref_cfg = to_env_config(d)                       # the reference's own loader on the same merged dict
        mine = kcfg.to_env_config(d)
        # the config mirror resolves to the same values
        for f in ("action_delta_scale", "episode_length", "dwell_steps_target", "dock_dynamic_residual_action_limit_near"):
            assert getattr(ref_cfg, f) == getattr(mine, f)
        for f in mine.reward_config.__dataclass_fields__:
            assert getattr(ref_cfg.reward_config, f) == getattr(mine.reward_config, f), f
        for f in mine.dock_reward_config.__dataclass_fields__:
            assert getattr(ref_cfg.dock_reward_config, f) == getattr(mine.dock_reward_config, f), f
        params = ko.params_from_config(mine)
        for ep in range(6):
            g = rng.uniform(-0.6, 0.6, 7); g[0] *= 0.3
            q0 = g + rng.uniform(-0.05, 0.05, 7) * (0.1 if mode == "dock" else 1.0)
            env = ref.ArmKinematicEnv(ref_cfg)
            obs, info = env.reset(options={"initial_q": q0, "goal_q": g, "policy_mode": mode})
            o = ko.OracleEnv(params)
            oobs = o.reset(mode=mode_idx, initial_q=q0, goal_q=g)
            keys = sorted(obs)
            assert np.array_equal(oobs, np.concatenate([obs[k] for k in keys]))
            for t in range(40):
                a = rng.uniform(-1.2, 1.2, 7) * rng.choice([0.05, 0.5, 1.0])
                obs, r, te, tr, info = env.step(a)
                oobs, out = o.step(a)
                assert abs(out.reward - r) < 1e-11 * max(1, abs(r))
                assert (bool(out.terminated), bool(out.truncated), bool(out.success)) == (te, tr, info["success"])
                assert np.abs(np.array(o.state.q[:]) - info["q"]).max() < 1e-13
                assert np.abs(oobs - np.concatenate([obs[k] for k in keys])).max() <= 1.2e-7
                if te or tr:
                    break
