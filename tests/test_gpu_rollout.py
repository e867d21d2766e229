"""GPU parity: policy forward, fused Approach -> Finisher rollout, samplers / auto-reset, the 1-env adapter."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import kin_oracle as ko

from ._util import env_config, golden, oracle_params, oracle_policy

pytestmark = pytest.mark.gpu


def _policy(name):
    from rl_brain_trainer_b200.policy import PolicyWeights

    return PolicyWeights.preset(name, "cuda")


def test_policy_forward_matches_oracle():
    rng = np.random.default_rng(1)
    for name, dim in (("approach_stage8_11", 56), ("finisher", 56), ("route_prefix120", 80)):
        pol, orc = _policy(name), oracle_policy(name)
        for n in (1, 127, 300):
            x = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
            a, v = pol.predict(torch.as_tensor(x), with_value=True)
            a, v = a.cpu().numpy(), v.cpu().numpy()
            for i in range(0, n, max(n // 16, 1)):
                ra, rv = orc.forward(x[i])
                assert np.abs(a[i] - ra).max() < 3e-6
                assert abs(v[i] - rv) < 3e-5 * max(1.0, abs(rv))


def _rollout(variant=0, approach="approach_dynamic_scale_big", policy="approach_stage8_11"):
    from rl_brain_trainer_b200.rollout import ApproachFinisherRollout

    return ApproachFinisherRollout(env_config(approach), _policy(policy), env_config("finisher_noop_ft"), _policy("finisher"), variant=variant)


def _compare_eval(res, ref, n, max_flip_frac):
    flips = int(np.sum(res["success"].astype(int) != ref["success"].astype(int)))
    assert flips <= max_flip_frac * n, f"{flips} of {n} success flags differ"
    same = res["success"].astype(int) == ref["success"].astype(int)
    same &= res["handoff_kind"] == ref["handoff_kind"]
    # where the discrete path agrees the continuous results agree to fp32 accuracy
    assert same.mean() > 1 - 2 * max_flip_frac - 0.02
    for name, tol in (("final_position_error", 2e-5), ("final_orientation_error", 1e-4), ("approach_final_position_error", 2e-5),
                      ("min_position_error", 2e-5)):
        d = np.abs(res[name][same] - ref[name][same])
        assert np.quantile(d, 0.99) < tol, (name, float(d.max()))
    assert np.array_equal(res["approach_steps"], ref["approach_steps"])
    return flips


def test_fused_rollout_stage5_golden():
    """Same 64 episodes the reference itself was run on (tests/golden/eval_stage5.npz): 61/64 successes."""
    from rl_brain_trainer_b200.samplers import EvalSuite

    g = golden("eval_stage5.npz")
    ro = _rollout()
    res = ro.evaluate_suite(EvalSuite(initial_q=g["initial_q"], goal_q=g["goal_q"], goal_pose6=g["goal_pose6"])).to_numpy()
    ref = {k: g[k] for k in ("success", "handoff_kind", "final_position_error", "final_orientation_error", "approach_final_position_error",
                             "min_position_error", "approach_steps")}
    flips = _compare_eval(res, ref, 64, 0.04)
    assert abs(res["success"].mean() - 0.953125) <= 2 / 64 + 1e-9, (res["success"].mean(), flips)
    assert abs(res["final_position_error"].mean() - g["final_position_error"].mean()) < 5e-5
    assert abs(res["final_orientation_error"].mean() - g["final_orientation_error"].mean()) < 5e-4


def test_fused_rollout_matches_oracle_4096():
    from rl_brain_trainer_b200.samplers import build_curriculum_local_eval_suite

    cfg = env_config("approach_dynamic_scale_big")
    suite = build_curriculum_local_eval_suite(cfg, seed=700001 + 5 * 1009, stage_index=5, n_episodes=4096)
    ro = _rollout()
    out = ro.evaluate_suite(suite)
    res = out.to_numpy()
    pa, pf = oracle_params(cfg), oracle_params(env_config("finisher_noop_ft"))
    ref, steps = ko.eval_approach_finisher(pa, pf, oracle_policy("approach_stage8_11"), oracle_policy("finisher"),
                                           initial_q=suite.initial_q.astype(np.float32).astype(float),
                                           goal_q=suite.goal_q.astype(np.float32).astype(float), n_threads=8)
    _compare_eval(res, ref, 4096, 0.01)
    assert abs(res["success"].mean() - ref["success"].mean()) < 0.01
    assert int(out.env_steps.item()) == int(res["approach_steps"].sum() + res["finisher_steps"].sum())
    assert abs(int(out.env_steps.item()) - steps) <= 36 * 0.02 * 4096


def test_fused_rollout_randomstart_known_split():
    """Mixed random-start known-workspace split, seed 940001: the reference gets 77/96 = 0.802."""
    from rl_brain_trainer_b200 import workspace

    cfg = env_config("randomstart_overnight")
    suites = workspace.build_randomstart_eval(cfg, seed=940001)
    g = golden("eval_randomstart.npz")
    ro = _rollout(approach="randomstart_overnight", policy="randomstart")
    for split, expect in (("known", 77), ("frontier", 23), ("stress", 20)):
        res = ro.evaluate_suite(suites[split]).to_numpy()
        ref = {k: g[f"{split}_{k}"] for k in ("success", "handoff_kind", "final_position_error", "final_orientation_error",
                                              "approach_final_position_error", "min_position_error", "approach_steps")}
        _compare_eval(res, ref, 96, 0.05)
        assert abs(int(res["success"].sum()) - expect) <= 3


def test_device_sampler_distribution_and_autoreset():
    from rl_brain_trainer_b200 import samplers
    from rl_brain_trainer_b200.env import BatchedArmKinematicEnv

    for preset, stage in (("approach_dynamic_scale_big", 9), ("randomstart_overnight", 10)):
        cfg = env_config(preset)
        n = 32768
        env = BatchedArmKinematicEnv(cfg, n, "cuda", seed=123, host_sampler=False)
        env.set_curriculum_stage(stage)
        env.reset()
        q, g = env.q.cpu().numpy(), env.goal_q.cpu().numpy()
        rng = np.random.default_rng(0)
        ref = [samplers.sample_reset(rng, cfg, "approach", stage) for _ in range(8192)]
        rq, rg = np.array([r.initial_q for r in ref]), np.array([r.goal_q for r in ref])
        for dev, host in ((q, rq), (g, rg)):
            assert np.abs(dev.mean(0) - host.mean(0)).max() < 0.04
            assert np.abs(dev.std(0) - host.std(0)).max() < 0.04
            assert np.abs(np.quantile(dev, [0.05, 0.5, 0.95], axis=0) - np.quantile(host, [0.05, 0.5, 0.95], axis=0)).max() < 0.08
        # goal pose == FK(goal_q), ee pose == FK(q)
        assert torch.allclose(env.goal_pose6, env.fk_pose6(env.goal_q), atol=1e-6)
        assert torch.allclose(env.ee_pose6, env.fk_pose6(env.q), atol=1e-6)
        # a different seed gives a different draw, the same seed the same one
        env2 = BatchedArmKinematicEnv(cfg, n, "cuda", seed=123, host_sampler=False)
        env2.set_curriculum_stage(stage)
        env2.reset()
        assert torch.equal(env2.q, env.q)

    cfg = env_config("approach_dynamic_scale_big")
    n = 4096
    env = BatchedArmKinematicEnv(cfg, n, "cuda", seed=7, host_sampler=False, auto_reset=True)
    env.set_curriculum_stage(5)
    env.reset()
    plain = BatchedArmKinematicEnv(cfg, n, "cuda", seed=7, host_sampler=False)
    plain.set_curriculum_stage(5)
    plain.reset()
    assert torch.equal(plain.state, env.state)
    T = cfg.termination_config.max_episode_steps
    a = torch.zeros(n, 7, device="cuda")
    for t in range(T):
        obs, r, te, tr, info = env.step(a)
        pobs, pr, pte, ptr, pinfo = plain.step(a)
        if t < T - 1:
            assert not bool(tr.any()) and torch.equal(obs, pobs) and torch.equal(r, pr)
    assert bool(tr.all()) and bool(info["auto_reset"].all())
    assert torch.equal(info["terminal_observation"], pobs)         # last obs of the finished episode
    assert torch.equal(r, pr)
    assert int(info["step_count"].max()) == 0                      # fresh episode
    assert not torch.equal(env.q, plain.q)
    obs2, *_ = env.step(a)
    assert int(env.counters()["step_count"].min()) == 1


def test_single_env_adapter_conformance():
    """The reference's own env tests (TESTS/test_kinematic_phase1_env.py), run against the GPU adapter."""
    from rl_brain_trainer_b200.config import Phase1EnvConfig
    from rl_brain_trainer_b200.env import ArmKinematicEnv

    env = ArmKinematicEnv(Phase1EnvConfig())
    obs, info = env.reset(seed=0)
    assert set(obs) == {"q", "dq", "prev_action", "goal_pos_err", "goal_ori_err", "wp_pos_err", "wp_ori_err", "next_wp_pos_err",
                        "next_wp_ori_err", "task_type", "mode_flag", "progress", "joint_limit_margin"}
    assert obs["q"].shape == (7,) and obs["mode_flag"].shape == (4,) and obs["q"].dtype == np.float32
    assert info["reason"] == "reset" and info["curriculum_stage_name"] == "region_small"
    # clipped state after an out-of-range action
    obs, reward, terminated, truncated, info = env.step(np.full(7, 10.0))
    assert np.all(np.abs(obs["prev_action"]) <= 1.0) and isinstance(reward, float) and isinstance(terminated, bool)
    lo = np.array([s.lower for s in env.config.joint_specs]); hi = np.array([s.upper for s in env.config.joint_specs])
    assert np.all(info["q"] >= lo - 1e-6) and np.all(info["q"] <= hi + 1e-6)
    with pytest.raises(ValueError):
        env.step(np.zeros(6))
    # success after success_dwell_steps zero-actions when reset at the goal
    g = np.array([0.0, 0.1, -0.1, 0.05, 0.0, 0.0, 0.0])
    env.reset(options={"initial_q": g, "goal_q": g})
    flags = [env.step(np.zeros(7)) for _ in range(2)]
    assert flags[0][4]["success"] is False and flags[1][4]["success"] is True and flags[1][2] is True and flags[1][4]["reason"] == "success"
    # stage reset uses FK(goal_q)
    env.set_curriculum_stage(3)
    _, info = env.reset(seed=5)
    assert np.abs(ko.fk_pose6(info["goal_q"]) - info["goal_pose6"]).max() < 1e-5
    assert env.get_curriculum_stage() == 3
    assert set(info) >= {"position_error_norm", "orientation_error_norm", "q", "dq", "goal_q", "goal_pose6", "success", "min_position_error",
                         "dwell_count", "near_goal_hit", "pre_near_goal_hit"}
    _, _, _, _, info = env.step(np.zeros(7))
    assert {"reward_components", "action_l2", "executed_delta_q_l2", "delta_q_change_l2", "dock_action_limit"} <= set(info)
    assert "position_progress" in info["reward_components"]


def test_single_env_adapter_seeded_reset_matches_reference_stream():
    from rl_brain_trainer_b200.env import ArmKinematicEnv

    g = golden("samplers.npz")
    cfg = env_config("approach_dynamic_scale_big")
    env = ArmKinematicEnv(cfg)
    env.set_curriculum_stage(5)
    env.reset(seed=1005)
    for i in range(8):
        _, info = env.reset()
        assert np.abs(info["q"] - g["approach_s5_q"][i]).max() < 1e-7
        assert np.abs(info["goal_q"] - g["approach_s5_goal_q"][i]).max() < 1e-7
        assert np.abs(info["goal_pose6"] - g["approach_s5_goal_pose6"][i]).max() < 1e-5


# ---- tensor-core variant (tcgen05 kind::tf32): statistical parity, the strict path is the FFMA variant above -------------------
def test_fused_rollout_tc_close_to_strict_fp32():
    from rl_brain_trainer_b200.rollout import VARIANT_FFMA, VARIANT_TC
    from rl_brain_trainer_b200.samplers import build_curriculum_local_eval_suite

    cfg = env_config("approach_dynamic_scale_big")
    for n in (100, 8192):   # ragged (partial tile / partial CTA) and multi-CTA sizes
        suite = build_curriculum_local_eval_suite(cfg, seed=700001 + 5 * 1009, stage_index=5, n_episodes=n)
        a = _rollout(VARIANT_FFMA).evaluate_suite(suite).to_numpy()
        b = _rollout(VARIANT_TC).evaluate_suite(suite).to_numpy()
        flips = int((a["success"] != b["success"]).sum())
        assert flips <= max(1, 0.005 * n), f"{flips} of {n} success flags differ between tf32 and fp32 MLP"
        assert abs(a["success"].mean() - b["success"].mean()) <= max(0.005, 1.0 / n) + 1e-9
        assert np.array_equal(a["approach_steps"], b["approach_steps"])
        assert float(np.abs(a["final_q"] - b["final_q"]).mean()) < 5e-4      # TF32 operands: O(1e-3) action differences
        assert abs(a["final_position_error"].mean() - b["final_position_error"].mean()) < 2e-5
        assert abs(a["final_orientation_error"].mean() - b["final_orientation_error"].mean()) < 2e-4


def test_fused_rollout_tc_reference_anchors():
    """Stage 5 (61/64) and random-start known split (77/96) with the tensor-core MLP: within +-3 episodes."""
    from rl_brain_trainer_b200 import workspace
    from rl_brain_trainer_b200.rollout import VARIANT_TC
    from rl_brain_trainer_b200.samplers import EvalSuite

    g = golden("eval_stage5.npz")
    res = _rollout(VARIANT_TC).evaluate_suite(EvalSuite(initial_q=g["initial_q"], goal_q=g["goal_q"], goal_pose6=g["goal_pose6"])).to_numpy()
    assert abs(int(res["success"].sum()) - 61) <= 3
    assert abs(res["final_position_error"].mean() - g["final_position_error"].mean()) < 1e-4
    suites = workspace.build_randomstart_eval(env_config("randomstart_overnight"), seed=940001)
    res = _rollout(VARIANT_TC, approach="randomstart_overnight", policy="randomstart").evaluate_suite(suites["known"]).to_numpy()
    assert abs(int(res["success"].sum()) - 77) <= 4


def test_fused_rollout_terminate_on_success_early_exit():
    """A config that terminates on success: lanes finish at different steps (per-lane predication, tile-uniform loop exit)."""
    import json
    from dataclasses import replace

    from rl_brain_trainer_b200.rollout import VARIANT_FFMA, VARIANT_TC, ApproachFinisherRollout
    from rl_brain_trainer_b200.samplers import build_curriculum_local_eval_suite

    cfg = env_config("approach_dynamic_scale_big")
    cfg_t = replace(cfg, termination_config=replace(cfg.termination_config, terminate_on_success=True))
    suite = build_curriculum_local_eval_suite(cfg, seed=11, stage_index=3, n_episodes=1000)
    pa, pf = oracle_params(cfg_t), oracle_params(env_config("finisher_noop_ft"))
    ref, _ = ko.eval_approach_finisher(pa, pf, oracle_policy("approach_stage8_11"), oracle_policy("finisher"),
                                       initial_q=suite.initial_q.astype(np.float32).astype(float),
                                       goal_q=suite.goal_q.astype(np.float32).astype(float), n_threads=8)
    assert ref["approach_steps"].min() < 128 and len(set(ref["approach_steps"].tolist())) > 5
    for variant, tol in ((VARIANT_FFMA, 0.01), (VARIANT_TC, 0.03)):
        ro = ApproachFinisherRollout(cfg_t, _policy("approach_stage8_11"), env_config("finisher_noop_ft"), _policy("finisher"), variant=variant)
        res = ro.evaluate_suite(suite).to_numpy()
        assert np.mean(res["approach_steps"] != ref["approach_steps"]) <= tol * 3
        assert np.mean(res["success"].astype(int) != ref["success"]) <= tol
    assert json is not None
