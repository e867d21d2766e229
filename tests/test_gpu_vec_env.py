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


def test_vec_env_against_the_oracle_through_an_autoreset_action_tape():
    """KinVecEnv driven next to the fp64 oracle by the same action tape over several auto-resets (SB3's VecEnv semantics): every
    env's reward / done / success / reason / position error and the observation SB3 would see agree with the oracle stepping the same
    episode; at a reset the oracle env is re-seated on the state the device sampler drew (initial q / dq / prev_action, goal), which is
    read back from the GPU env -- the sampler itself is tested distributionally elsewhere."""
    import ctypes

    from oracle import kin_oracle as ko
    from rl_brain_trainer_b200.env import OBS_KEYS, OBS_SLICES
    from rl_brain_trainer_b200.vec_env import make_vec_env

    from ._util import oracle_params

    cfg = env_config("approach_dynamic_scale_big")
    cfg = dataclasses.replace(cfg, episode_length=9, termination_config=dataclasses.replace(cfg.termination_config, max_episode_steps=9))
    n = 64
    venv = make_vec_env(cfg, n, seed=11, stage_index=3)
    params = oracle_params(cfg)
    states = ko.state_array(n)
    L = ko.lib()

    def reseat(envs):
        q, dq, pa = venv.env.q.cpu().numpy().astype(float), venv.env.dq.cpu().numpy().astype(float), venv.env.prev_action.cpu().numpy().astype(float)
        gq, gp = venv.env.goal_q.cpu().numpy().astype(float), venv.env.goal_pose6.cpu().numpy().astype(float)
        for e in envs:
            L.kor_reset(ctypes.byref(params), ctypes.byref(states[e]), 0, ko._dptr(q[e].copy()), ko._dptr(dq[e].copy()), ko._dptr(pa[e].copy()),
                        ko._dptr(gq[e].copy()), ko._dptr(gp[e].copy()))

    obs = venv.reset()
    reseat(range(n))
    order = sorted(OBS_KEYS, key=lambda k: OBS_SLICES[k].start)        # SB3's alphabetical flattening = the oracle's 56-vector
    rng = np.random.default_rng(4)
    episodes = 0
    for t in range(31):                               # three time limits per env + a tail
        a = (rng.uniform(-1, 1, (n, 7)) * (0.2 if t % 2 else 1.0)).astype(np.float32)
        obs, rew, dones, infos = venv.step(a)
        robs, routs = ko.step_batch(params, states, a.astype(float))
        r_rew = np.array([o.reward for o in routs])
        r_done = np.array([bool(o.terminated or o.truncated) for o in routs])
        assert np.array_equal(dones, r_done), t
        assert np.abs(rew - r_rew).max() < 2e-4 * max(1.0, np.abs(r_rew).max()), t
        for e in range(n):
            assert infos[e]["success"] == bool(routs[e].success) and infos[e]["TimeLimit.truncated"] == bool(routs[e].truncated and not routs[e].terminated)
            assert abs(infos[e]["position_error_norm"] - routs[e].position_error_norm) < 1e-5
        flat = np.concatenate([obs[k] for k in order], axis=1)
        live = ~dones
        assert not live.any() or np.abs(flat[live] - robs[live]).max() < 5e-5, t
        for e in np.nonzero(dones)[0]:                # SB3: the finished episode's last observation travels in infos
            term = np.concatenate([infos[e]["terminal_observation"][k] for k in order])
            assert np.abs(term - robs[e]).max() < 5e-5, (t, e)
        if dones.any():
            episodes += int(dones.sum())
            reseat(np.nonzero(dones)[0])              # the next episode starts from the state the device sampler drew
            q_obs = flat[dones][:, OBS_SLICES["progress"]]
            assert np.allclose(q_obs[:, 0], 0.0)      # ... and the observation SB3 sees is that episode's first one
    assert episodes >= 3 * n
    venv.close()
