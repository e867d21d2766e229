"""KinVecEnv: stable-baselines3's VecEnv protocol (auto-reset, terminal_observation, TimeLimit.truncated, env_method) on the GPU env,
checked step by step against the single-env adapter that mirrors the reference's ArmKinematicEnv."""

from __future__ import annotations

import dataclasses

import numpy as np
import pytest
import torch

from ._util import env_config

pytestmark = pytest.mark.gpu


def test_vec_env_protocol_and_autoreset_semantics():
    from rl_brain_trainer_b200.env import OBS_KEYS, ArmKinematicEnv
    from rl_brain_trainer_b200.vec_env import KinVecEnv, make_vec_env

    cfg = env_config("approach_dynamic_scale_big")
    cfg = dataclasses.replace(cfg, episode_length=12, termination_config=dataclasses.replace(cfg.termination_config, max_episode_steps=12))
    n = 48
    venv = make_vec_env(cfg, n, seed=5, stage_index=2)
    assert isinstance(venv, KinVecEnv) and venv.num_envs == n and set(venv.observation_space.spaces) == set(OBS_KEYS)
    obs = venv.reset()
    assert set(obs) == set(OBS_KEYS) and obs["q"].shape == (n, 7) and obs["progress"].shape == (n, 3) and obs["q"].dtype == np.float32
    assert venv.env_method("get_curriculum_stage") == [2] * n and venv.env_is_wrapped(object) == [False] * n
    # a single-env adapter started from env 3's state follows the same trajectory
    ref = ArmKinematicEnv(cfg)
    q0, gq = venv.env.q[3].cpu().numpy(), venv.env.goal_q[3].cpu().numpy()
    robs, _ = ref.reset(options={"initial_q": q0, "goal_q": gq})
    for k in OBS_KEYS:
        assert np.allclose(robs[k], obs[k][3], atol=1e-6), k
    rng = np.random.default_rng(0)
    for t in range(12):
        a = rng.uniform(-1, 1, (n, 7)).astype(np.float32)
        venv.step_async(a)
        obs, rew, dones, infos = venv.step_wait()
        robs, rrew, rterm, rtrunc, rinfo = ref.step(a[3])
        assert rew.dtype == np.float32 and dones.dtype == bool and len(infos) == n
        assert abs(float(rew[3]) - rrew) < 1e-5 and bool(dones[3]) == bool(rterm or rtrunc)
        assert infos[3]["success"] == bool(rinfo["success"]) and infos[3]["reason"] == rinfo["reason"]
        assert abs(infos[3]["position_error_norm"] - rinfo["position_error_norm"]) < 1e-6
        if t < 11:
            assert not dones.any() and "terminal_observation" not in infos[3] and not infos[3]["TimeLimit.truncated"]
            for k in OBS_KEYS:
                assert np.allclose(robs[k], obs[k][3], atol=1e-6), (t, k)
    # step 12 hits the time limit everywhere: SB3 semantics -- obs is the NEXT episode's first observation, the last one is in infos
    assert dones.all() and all(i["TimeLimit.truncated"] and i["reason"] == "max_steps" for i in infos)
    for k in OBS_KEYS:
        assert np.allclose(infos[3]["terminal_observation"][k], robs[k], atol=1e-6), k
    assert np.allclose(obs["progress"][:, 0], 0.0) and not np.allclose(obs["q"][3], robs["q"])
    # curriculum promotion through env_method reaches the device sampler at the next auto-reset
    venv.env_method("set_curriculum_stage", 7)
    for _ in range(12):
        obs, rew, dones, infos = venv.step(np.zeros((n, 7), dtype=np.float32))
    assert dones.all() and venv.env_method("get_curriculum_stage") == [7] * n
    # error behaviour
    with pytest.raises(ValueError):
        venv.step_async(np.zeros((n, 6), dtype=np.float32))
    with pytest.raises(AttributeError):
        venv.env_method("no_such_method")
    with pytest.raises(RuntimeError):
        venv.step_wait()
    venv.set_attr("tag", 3)
    assert venv.get_attr("tag") == [3] * n and venv.seed(9)[1] == 10
    venv.close()
