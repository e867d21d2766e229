"""GPU tests of the handoff-state buffer builder (vs outputs of the live reference) and of the device-side dock reset branches
(handoff-state replay, close-bucket rejection sampling) vs the host port of the reference sampler."""

from __future__ import annotations

import dataclasses
import json

import numpy as np
import pytest
import torch

from ._util import env_config, golden

pytestmark = pytest.mark.gpu


def _approach():
    from rl_brain_trainer_b200.policy import PolicyWeights

    return env_config("approach_dynamic_scale_big"), PolicyWeights.preset("approach_stage8_11", "cuda")


def test_handoff_buffer_matches_reference_builder(tmp_path):
    """48 Stage-10 episodes the reference's builder loop was run on (tests/golden/gen_golden_handoff.py): all three handoff modes."""
    from rl_brain_trainer_b200 import handoff
    from rl_brain_trainer_b200.samplers import EvalSuite

    g = golden("handoff_states.npz")
    cfg, pol = _approach()
    suite = EvalSuite(initial_q=g["initial_q"], goal_q=g["goal_q"], goal_pose6=g["goal_pose6"])
    n = len(suite)
    for mode, stored_ref, pre in (("final_settled", g["final_ready"] == 1, "final"), ("first_confirmed", g["has_first"] == 1, "first"),
                                  ("final_always", np.ones(n, dtype=bool), "final")):
        buf, summary = handoff.build_finisher_handoff_state_buffer(cfg, pol, suite=suite, stage_index=10, handoff_mode=mode,
                                                                   source_checkpoint_name="model_best_by_gate.zip")
        stored = np.zeros(n, dtype=bool)
        stored[buf.episode_id] = True
        # fp32 vs the reference's fp64: an episode exactly at a ready threshold may flip; none does on this draw (band kept for safety)
        assert int(np.sum(stored != stored_ref)) <= 1, mode
        both = stored & stored_ref
        sel = both[buf.episode_id]
        ids = buf.episode_id[sel]
        assert np.array_equal(buf.step_index[sel], g[f"{pre}_step"][ids])
        for name, key, tol in (("initial_q", "q", 2e-5), ("initial_dq", "dq", 2e-5), ("initial_prev_action", "prev_action", 2e-4),
                               ("goal_q", "goal_q", 1e-6), ("goal_pose6", "goal_pose6", 1e-6)):
            assert np.abs(getattr(buf, name)[sel] - g[f"{pre}_{key}"][ids]).max() < tol, (mode, name)
        for name, key in (("position_error_norm", "final_position_error"), ("orientation_error_norm", "final_orientation_error"),
                          ("action_l2", "final_action_magnitude"), ("dq_norm", "final_dq_norm")):
            assert np.abs(getattr(buf, name)[sel] - g[f"{pre}_{key}"][ids]).max() < 2e-4, (mode, name)
        assert summary["episode_count"] == n and summary["stored_handoff_count"] == len(buf) and summary["handoff_mode"] == mode
        assert len(summary["episode_summaries"]) == n and summary["states"][0]["dwell_count"] == int(g["dwell_steps_target"])
    # the JSON round trip through the reference's reader semantics (filters included)
    path = handoff.write_handoff_buffer(tmp_path, summary)
    assert path.name == "finisher_handoff_state_buffer.json" and set(json.loads(path.read_text())["states"][0]) >= {
        "episode_id", "step_index", "initial_q", "initial_dq", "initial_prev_action", "goal_q", "goal_pose6", "position_error_norm",
        "orientation_error_norm", "dwell_count", "action_l2", "dq_norm", "source_checkpoint_name", "handoff_mode"}
    back = handoff.load_handoff_states(path)
    assert len(back) == len(buf) and np.allclose(back.rows(), buf.rows())
    tight = handoff.load_handoff_states(path, max_position_error_m=float(np.median(buf.position_error_norm)))
    assert 0 < len(tight) < len(buf)
    with pytest.raises(FileNotFoundError):
        handoff.load_handoff_states(tmp_path / "missing.json")


def test_handoff_builder_agrees_with_the_fused_rollout_at_scale():
    """8 192 Stage-8 episodes: the builder's approach end state is the approach-only fused rollout's final state, bit for bit."""
    from rl_brain_trainer_b200 import handoff
    from rl_brain_trainer_b200.rollout import ApproachFinisherRollout
    from rl_brain_trainer_b200.samplers import build_curriculum_local_eval_suite

    cfg, pol = _approach()
    suite = build_curriculum_local_eval_suite(cfg, seed=700001, stage_index=8, n_episodes=8192)
    buf, summary = handoff.build_finisher_handoff_state_buffer(cfg, pol, suite=suite, stage_index=8, handoff_mode="final_always")
    res = ApproachFinisherRollout(cfg, pol).evaluate_suite(suite)
    assert len(buf) == 8192 and np.array_equal(buf.initial_q.astype(np.float32), res.final_q.cpu().numpy())
    assert np.array_equal(buf.position_error_norm.astype(np.float32), res.final_position_error.cpu().numpy())
    assert np.array_equal(buf.step_index, res.approach_steps.cpu().numpy().astype(int))
    settled, _ = handoff.build_finisher_handoff_state_buffer(cfg, pol, suite=suite, stage_index=8, handoff_mode="final_settled")
    assert np.array_equal(np.sort(settled.episode_id), np.nonzero(res.final_ready.cpu().numpy())[0])
    assert 0.3 < summary["stored_handoff_rate"] <= 1.0 and summary["env_steps"] == int(res.env_steps.item())


def _dock_env(dock_reset, n=8192, seed=3):
    from rl_brain_trainer_b200.env import BatchedArmKinematicEnv

    cfg = env_config("finisher_noop_ft")
    cfg = dataclasses.replace(cfg, dock_reset_config=dataclasses.replace(cfg.dock_reset_config, **dock_reset))
    return cfg, BatchedArmKinematicEnv(cfg, n, "cuda", auto_reset=True, seed=seed, host_sampler=False, with_aux=False)


def test_device_dock_reset_replays_handoff_states():
    """reset_samplers.py:434-446 on the device: with probability p a reset copies one buffer row (q, dq, prev_action, goal pose)."""
    from rl_brain_trainer_b200 import handoff
    from rl_brain_trainer_b200.samplers import build_curriculum_local_eval_suite

    acfg, pol = _approach()
    buf, _ = handoff.build_finisher_handoff_state_buffer(acfg, pol, suite=build_curriculum_local_eval_suite(acfg, seed=11, stage_index=6, n_episodes=512),
                                                         stage_index=6, handoff_mode="final_always")
    rows = buf.device_rows("cuda")
    p = 0.6
    cfg, env = _dock_env({"handoff_state_probability": p})
    env.set_handoff_states(rows)
    env.reset()
    torch.cuda.synchronize()
    n = env.num_envs
    state = torch.cat([env.q, env.dq, env._rows("KIN_ROW_PREV_ACTION", 7), env._rows("KIN_ROW_GOAL_POSE", 6)], dim=1)      # [n, 27]
    key = torch.cat([rows[:, :21], rows[:, 28:34]], dim=1)                                                                   # [M, 27]
    match = (torch.cdist(state.double(), key.double(), p=float("inf")).min(dim=1).values == 0.0)
    frac = float(match.float().mean())
    assert abs(frac - p) < 4 * np.sqrt(p * (1 - p) / n) + 0.005, frac
    # the replayed states use many different rows, the others are ordinary dock resets (start within init_q_noise of the goal)
    idx = torch.cdist(state[match].double(), key.double(), p=float("inf")).argmin(dim=1)
    assert int(torch.unique(idx).numel()) > 0.9 * min(len(buf), int(match.sum()) * 0.6)
    other = ~match
    gq = env._rows("KIN_ROW_GOAL_Q", 7)
    d = cfg.dock_reset_config       # the preset also uses the close-bucket branch: its noise box is the wider one
    noise = torch.tensor(np.maximum(d.init_q_noise, d.close_init_q_noise if d.close_bucket_probability > 0 else 0.0), device="cuda", dtype=torch.float32)
    assert bool(((env.q[other] - gq[other]).abs() <= noise + 1e-6).all()) and bool((env.dq[other] == 0).all())
    # auto-reset inside the step kernel draws from the same sampler
    for _ in range(cfg.termination_config.max_episode_steps + 1):
        env.step_raw(torch.zeros((n, 7), device="cuda"))
    torch.cuda.synchronize()
    state2 = torch.cat([env.q, env.dq, env._rows("KIN_ROW_PREV_ACTION", 7), env._rows("KIN_ROW_GOAL_POSE", 6)], dim=1)
    assert not torch.equal(state, state2)
    # clearing the buffer turns the branch off
    env.set_handoff_states(None)
    env.reset()
    state3 = torch.cat([env.q, env.dq, env._rows("KIN_ROW_PREV_ACTION", 7), env._rows("KIN_ROW_GOAL_POSE", 6)], dim=1)
    assert not bool((torch.cdist(state3.double(), key.double(), p=float("inf")).min(dim=1).values == 0.0).any())


def test_device_dock_reset_close_bucket_matches_host_port():
    """reset_samplers.py:452-515 on the device: starts land in the (position, orientation) error bucket as often as the host port's."""
    from rl_brain_trainer_b200 import samplers
    from rl_brain_trainer_b200.kinematics import fk_pose6_folded as fk_pose6_numpy

    bucket = {"close_bucket_probability": 1.0, "close_bucket_min_pos_error_m": 0.006, "close_bucket_max_pos_error_m": 0.012,
              "close_bucket_min_ori_error_rad": 0.0, "close_bucket_max_ori_error_rad": 0.05, "close_bucket_max_attempts": 6}
    cfg, env = _dock_env(bucket, n=16384)
    env.set_curriculum_stage(4)
    obs, info = env.reset()
    torch.cuda.synchronize()
    pos, ori = info["position_error_norm"].cpu().numpy(), info["orientation_error_norm"].cpu().numpy()
    inside = (pos >= 0.006 - 1e-6) & (pos <= 0.012 + 1e-6) & (ori <= 0.05 + 1e-6)
    # host port of the reference (PCG64), same config: the acceptance rate after <= 6 attempts is a property of the distribution
    rng = np.random.default_rng(0)
    host_inside = []
    for _ in range(1500):
        s = samplers.sample_dock_reset(rng, cfg, 4, fk=fk_pose6_numpy)
        gp, ip = fk_pose6_numpy(s.goal_q), fk_pose6_numpy(s.initial_q)
        pe = np.linalg.norm(gp[:3] - ip[:3])
        oe = np.linalg.norm((gp[3:] - ip[3:] + np.pi) % (2 * np.pi) - np.pi)
        host_inside.append(0.006 <= pe <= 0.012 and oe <= 0.05)
    h = float(np.mean(host_inside))
    assert 0.05 < h < 0.999         # the bucket is neither trivial nor unreachable within 6 attempts
    assert abs(float(inside.mean()) - h) < 4 * np.sqrt(h * (1 - h) / 1500) + 0.01, (float(inside.mean()), h)
    # misses keep the candidate closest to the bucket: still near the goal
    assert float(pos.max()) < 0.05
    # probability 0 -> the plain branch (no bucket structure)
    _, env0 = _dock_env({"close_bucket_probability": 0.0}, n=4096)
    env0.set_curriculum_stage(4)
    _, info0 = env0.reset()
    pos0 = info0["position_error_norm"].cpu().numpy()
    assert abs(float(((pos0 >= 0.006) & (pos0 <= 0.012)).mean()) - float(inside.mean())) > 0.1
