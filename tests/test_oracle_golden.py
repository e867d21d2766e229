"""Pin the CPU fp64 oracle (oracle/kin_oracle.c) against golden vectors recorded from the LIVE reference.

Fixtures: tests/golden/*.npz, written by tests/golden/gen_golden.py (which imports the reference from
/root/reference).  Tolerances are fp64 round-off (<= 1e-12 relative); flags and counters are exact.
"""

from __future__ import annotations

import numpy as np
import pytest

from oracle import kin_oracle as ko

from ._util import GOLD, env_config, env_config_from_json, golden, oracle_params, oracle_policy, policy_weights

TOL = 1e-11


def test_fk_pose_and_matrix():
    g = golden("fk.npz")
    pose = ko.fk_pose6(g["q"])
    assert np.abs(pose - g["pose6"]).max() < 1e-12
    for i in range(g["mats"].shape[0]):
        assert np.abs(ko.fk_matrix(g["q"][i]) - g["mats"][i]).max() < 1e-13
    # home pose quoted in SURVEY 8(c)
    home = ko.fk_pose6(np.zeros(7))
    assert np.allclose(home, [-0.180010258, 0.00176389303, 1.10045, 1.57079633, 6.72e-15, -1.57079633], atol=1e-8)


def test_wrap_to_pi():
    g = golden("fk.npz")
    out = np.array([ko.wrap_to_pi(v) for v in g["wrap_in"]])
    assert np.abs(out - g["wrap_out"]).max() < 1e-14


def _replay(trace, params, names):
    """Replay every recorded episode through the oracle and compare every recorded output."""
    starts = trace["episode_start"]
    env = ko.OracleEnv(params)
    n_checked = 0
    for e in range(len(starts) - 1):
        gp = trace["reset_goal_pose6"][e] if trace["reset_has_goal_pose6"][e] else None
        obs = env.reset(mode=int(trace["reset_mode"][e]), initial_q=trace["reset_initial_q"][e], goal_q=trace["reset_goal_q"][e],
                        goal_pose6=gp, initial_dq=trace["reset_initial_dq"][e], initial_prev_action=trace["reset_initial_prev_action"][e])
        assert np.array_equal(obs, trace["reset_obs"][e]), f"reset obs episode {e}"
        s = env.state
        assert np.abs(np.array(s.ee_pose6[:]) - trace["reset_ee_pose6"][e]).max() < 1e-12
        entry = np.array([s.entry_position_error_norm, s.entry_orientation_error_norm, s.entry_action_l2, s.entry_dq_norm])
        assert np.abs(entry - trace["reset_entry"][e]).max() < 1e-12
        for t in range(starts[e], starts[e + 1]):
            obs, out = env.step(trace["action"][t])
            where = f"episode {e} step {t - starts[e]}"
            assert abs(out.reward - trace["reward"][t]) <= TOL * max(1.0, abs(trace["reward"][t])), where
            comps = np.array(out.components[: len(names)])
            ref = trace["components"][t][: len(names)]
            bad = np.abs(comps - ref) > TOL * np.maximum(1.0, np.abs(ref))
            assert not bad.any(), f"{where}: components {[names[i] for i in np.nonzero(bad)[0]]}"
            assert (out.terminated, out.truncated, out.success, out.reason) == (
                trace["terminated"][t], trace["truncated"][t], trace["success"][t], trace["reason"][t]), where
            assert np.abs(np.array(s.q[:]) - trace["q"][t]).max() < 1e-13, where
            assert np.abs(np.array(s.dq[:]) - trace["dq"][t]).max() < 1e-13, where
            assert np.abs(np.array(s.prev_action[:]) - trace["prev_action"][t]).max() < 1e-13, where
            assert np.abs(np.array(s.ee_pose6[:]) - trace["ee_pose6"][t]).max() < 1e-12, where
            assert abs(out.position_error_norm - trace["pos_err"][t]) < 1e-13, where
            assert abs(out.orientation_error_norm - trace["ori_err"][t]) < 1e-12, where
            assert abs(out.action_l2 - trace["action_l2"][t]) < 1e-13, where
            assert abs(out.executed_delta_q_l2 - trace["dq_l2"][t]) < 1e-14, where
            assert abs(out.delta_q_change_l2 - trace["dq_change_l2"][t]) < 1e-14, where
            assert abs(out.dock_action_limit - trace["dock_action_limit"][t]) < 1e-13, where
            assert abs(s.min_pos_error - trace["min_pos_err"][t]) < 1e-13, where
            assert abs(out.joint_limit_margin_min - trace["margin_min"][t]) < 1e-14, where
            counters = np.array([s.episode_step, s.dwell_count, s.near_goal_entry_count, s.near_goal_drift_count,
                                 s.pre_near_goal_hit, s.near_goal_hit, out.curr_in_pre_near_goal, out.curr_in_near_goal])
            assert np.array_equal(counters, trace["counters"][t]), where
            # obs is fp32 of an fp64 value: a 1e-16 difference can flip the last float bit
            assert np.abs(obs - trace["obs"][t]).max() <= 1.2e-7, where
            n_checked += 1
    return n_checked


def test_step_trace_approach_official():
    n = _replay(golden("trace_approach.npz"), oracle_params(env_config("approach_dynamic_scale_big")), ko.APPROACH_COMPONENT_NAMES)
    assert n > 1500


def test_step_trace_dock_official():
    n = _replay(golden("trace_dock.npz"), oracle_params(env_config("finisher_noop_ft")), ko.DOCK_COMPONENT_NAMES)
    assert n > 400


def test_step_trace_dock_all_knobs():
    cfg = env_config_from_json(GOLD / "trace_dock_alt_config.json")
    assert _replay(golden("trace_dock_alt.npz"), oracle_params(cfg), ko.DOCK_COMPONENT_NAMES) > 40


def test_step_trace_approach_all_knobs():
    cfg = env_config_from_json(GOLD / "trace_approach_alt_config.json")
    assert _replay(golden("trace_approach_alt.npz"), oracle_params(cfg), ko.APPROACH_COMPONENT_NAMES) > 30


def test_reference_known_answers():
    """The handful of exact values the reference's own tests hold."""
    cfg = env_config("approach_default")
    p = oracle_params(cfg)
    env = ko.OracleEnv(p)
    # mode flags one-hot positions (TESTS/test_kinematic_phase1_split.py:21-35)
    for mode, idx in ((ko.MODE_APPROACH, 0), (ko.MODE_DOCK, 1)):
        obs = env.reset(mode=mode, initial_q=np.zeros(7), goal_q=np.zeros(7))
        assert list(obs[20:24]) == [1.0 if i == idx else 0.0 for i in range(4)]
        assert list(obs[47:50]) == [1.0, 0.0, 0.0]
    # success after success_dwell_steps zero actions when reset at the goal (test_kinematic_phase1_env.py:51-60)
    env.reset(mode=ko.MODE_APPROACH, initial_q=np.zeros(7), goal_q=np.zeros(7))
    _, o1 = env.step(np.zeros(7))
    assert not o1.success and o1.reason == 0
    _, o2 = env.step(np.zeros(7))
    assert o2.success and o2.terminated and ko.REASONS[o2.reason] == "success"
    # clipped state after an out-of-range action
    env.reset(mode=ko.MODE_APPROACH, initial_q=[0.38, 3.1, 0, 0, 0, 0, 0], goal_q=np.zeros(7))
    env.step(np.full(7, 5.0))
    assert env.state.q[0] <= 0.385 and env.state.q[1] <= np.pi and env.state.prev_action[0] == 1.0
    # re-entry bonus decay 1.0 -> 0.5 and leave penalty -0.35 (test_kinematic_phase1_approach_reward.py:119-120,186)
    from dataclasses import replace

    from rl_brain_trainer_b200 import config as kcfg

    base = kcfg.Phase1EnvConfig()
    cfg2 = replace(base, reward_config=replace(base.reward_config, near_goal_leave_penalty=0.35),
                   termination_config=replace(base.termination_config, terminate_on_success=False))
    env = ko.OracleEnv(oracle_params(cfg2))
    env.reset(mode=ko.MODE_APPROACH, initial_q=[0, 0, 0.24, 0, 0, 0, 0], goal_q=np.zeros(7))
    names = ko.APPROACH_COMPONENT_NAMES
    toward, away = np.array([0, 0, -1.0, 0, 0, 0, 0]), np.array([0, 0, 1.0, 0, 0, 0, 0])
    _, o = env.step(toward)
    assert o.components[names.index("near_goal_bonus")] == pytest.approx(0.10)
    assert o.components[names.index("near_goal_bonus_scale")] == 1.0
    _, o = env.step(away)
    assert o.components[names.index("near_goal_leave_penalty")] == pytest.approx(-0.35)
    _, o = env.step(toward)
    assert o.components[names.index("near_goal_bonus_scale")] == 0.5
    assert o.components[names.index("near_goal_bonus")] == pytest.approx(0.05)


def test_dynamic_dock_limits_endpoints():
    """Dynamic limit interpolation end values 0.11 / 0.04 (TESTS/test_kinematic_phase1_split.py:127-128)."""
    cfg = env_config_from_json(GOLD / "trace_dock_alt_config.json")
    p = oracle_params(cfg)
    env = ko.OracleEnv(p)
    g = np.array([0.05, 0.2, -0.3, 0.1, 0.2, -0.1, 0.3])
    env.reset(mode=ko.MODE_DOCK, initial_q=g, goal_q=g)
    _, out = env.step(np.ones(7))
    assert out.dock_action_limit == pytest.approx(0.11) and out.dock_delta_q_change_limit_scale == pytest.approx(0.04)
    assert max(abs(a) for a in env.state.prev_action[:]) == pytest.approx(0.11)


def test_termination_truth_table():
    """termination.py:20-57 with terminate_on_success False/True (test_kinematic_phase1_split.py:588-624)."""
    cfg = env_config("approach_dynamic_scale_big")
    p = oracle_params(cfg)
    assert not p.term_terminate_on_success
    env = ko.OracleEnv(p)
    g = np.array([0.0, 0.1, -0.1, 0.05, 0.0, 0.02, 0.0])
    env.reset(mode=ko.MODE_APPROACH, initial_q=g, goal_q=g)
    flags = []
    for _ in range(128):
        _, out = env.step(np.zeros(7))
        flags.append((out.success, out.terminated, out.truncated))
    assert flags[0] == (0, 0, 0) and flags[1] == (1, 0, 0) and flags[-1] == (1, 0, 1)
    assert all(not f[1] for f in flags) and sum(f[2] for f in flags) == 1


def test_policy_forward_matches_torch():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(3)
    for name, dim in (("approach_stage8_11", 56), ("finisher", 56), ("route_prefix120", 80)):
        w = policy_weights(name)
        pol = oracle_policy(name)
        x = rng.uniform(-1, 1, size=(16, dim)).astype(np.float32)
        t = {k: torch.from_numpy(v) for k, v in w.items()}
        lin = torch.nn.functional.linear
        h = torch.tanh(lin(torch.from_numpy(x), t["mlp_extractor.policy_net.0.weight"], t["mlp_extractor.policy_net.0.bias"]))
        h = torch.tanh(lin(h, t["mlp_extractor.policy_net.2.weight"], t["mlp_extractor.policy_net.2.bias"]))
        a_ref = lin(h, t["action_net.weight"], t["action_net.bias"]).clamp(-1, 1).numpy()
        hv = torch.tanh(lin(torch.from_numpy(x), t["mlp_extractor.value_net.0.weight"], t["mlp_extractor.value_net.0.bias"]))
        hv = torch.tanh(lin(hv, t["mlp_extractor.value_net.2.weight"], t["mlp_extractor.value_net.2.bias"]))
        v_ref = lin(hv, t["value_net.weight"], t["value_net.bias"]).numpy()[:, 0]
        for i in range(16):
            a, v = pol.forward(x[i])
            assert np.abs(a - a_ref[i]).max() < 2e-6
            assert abs(v - v_ref[i]) < 2e-5 * max(1.0, abs(v_ref[i]))


def _check_eval(res, g, prefix=""):
    """Closed-loop results: identical flags/steps, errors to 1e-6 (fp32 policy arithmetic order differs from torch)."""
    k = lambda name: g[prefix + name]  # noqa: E731
    flips = int(np.sum(res["success"] != k("success")))
    assert flips == 0, f"{flips} success flags differ"
    assert np.array_equal(res["handoff_kind"], k("handoff_kind"))
    assert np.array_equal(res["handoff_step"], k("handoff_step"))
    assert np.array_equal(res["first_ready_step"], k("first_ready_step"))
    assert np.array_equal(res["max_ready_streak"], k("max_ready_streak"))
    assert np.array_equal(res["approach_steps"], k("approach_steps"))
    assert np.array_equal(res["finisher_steps"], k("finisher_steps"))
    assert np.array_equal(res["ready_hit"], k("ready_hit")) and np.array_equal(res["ready_dwell"], k("ready_dwell"))
    for name, tol in (("final_position_error", 2e-6), ("final_orientation_error", 2e-5), ("approach_final_position_error", 2e-6),
                      ("approach_final_orientation_error", 2e-5), ("min_position_error", 2e-6), ("final_action_magnitude", 2e-4),
                      ("final_dq_norm", 2e-5)):
        assert np.abs(res[name] - k(name)).max() < tol, name
    assert np.abs(res["final_q"] - k("final_q")).max() < 1e-4


def test_eval_stage5_reproduces_reference():
    g = golden("eval_stage5.npz")
    pa, pf = oracle_params(env_config("approach_dynamic_scale_big")), oracle_params(env_config("finisher_noop_ft"))
    res, steps = ko.eval_approach_finisher(pa, pf, oracle_policy("approach_stage8_11"), oracle_policy("finisher"),
                                           initial_q=g["initial_q"], goal_q=g["goal_q"], goal_pose6=g["goal_pose6"], n_threads=4)
    _check_eval(res, g)
    assert steps == int(g["approach_steps"].sum() + g["finisher_steps"].sum())
    assert res["success"].mean() == pytest.approx(0.953125)  # 61/64 with the bundled checkpoints (published: 0.93)


def test_eval_other_stages():
    g = golden("eval_stages.npz")
    pa, pf = oracle_params(env_config("approach_dynamic_scale_big")), oracle_params(env_config("finisher_noop_ft"))
    for stage in (0, 8, 11):
        pre = f"s{stage}_"
        res, _ = ko.eval_approach_finisher(pa, pf, oracle_policy("approach_stage8_11"), oracle_policy("finisher"),
                                           initial_q=g[pre + "initial_q"], goal_q=g[pre + "goal_q"], goal_pose6=g[pre + "goal_pose6"])
        _check_eval(res, g, pre)


def test_eval_randomstart_reproduces_reference():
    """Mixed random-start eval, seed 940001: known-workspace split 77/96 = 0.802 (report/OFFICIAL_ARTIFACTS.md:168)."""
    g = golden("eval_randomstart.npz")
    pa, pf = oracle_params(env_config("randomstart_overnight")), oracle_params(env_config("finisher_noop_ft"))
    expect = {"known": 77, "frontier": 23, "stress": 20}
    for split in ("known", "frontier", "stress"):
        pre = split + "_"
        res, _ = ko.eval_approach_finisher(pa, pf, oracle_policy("randomstart"), oracle_policy("finisher"),
                                           initial_q=g[pre + "initial_q"], initial_dq=g[pre + "initial_dq"],
                                           initial_prev_action=g[pre + "initial_prev_action"], goal_q=g[pre + "goal_q"],
                                           goal_pose6=g[pre + "goal_pose6"], n_threads=4)
        _check_eval(res, g, pre)
        assert int(res["success"].sum()) == expect[split]


def _route_setup():
    from rl_brain_trainer_b200 import config as kcfg

    cfg_dict = kcfg.preset_dict("route_prefix120")
    renv, seq = kcfg.to_route_env_config(cfg_dict)
    return oracle_params(renv.base_env_config, renv.reward_config), renv, seq


def test_route_dataset_and_traces():
    g = golden("trace_route.npz")
    params, renv, _ = _route_setup()
    route = ko.OracleRoute(g["route_q"])
    assert np.abs(route.pose6 - g["route_pose6"]).max() < 1e-12
    assert np.abs(route.progress_m - g["route_progress"]).max() < 1e-12
    names = ko.ROUTE_COMPONENT_NAMES

    def check(prefix, t, obs, out):
        where = f"{prefix} step {t}"
        assert abs(out.route_reward - g[prefix + "reward"][t]) < 1e-11 * max(1, abs(g[prefix + "reward"][t])), where
        comps = np.array(out.route_components[:17])
        bad = np.abs(comps - g[prefix + "components"][t]) > 1e-11 * np.maximum(1, np.abs(g[prefix + "components"][t]))
        assert not bad.any(), f"{where}: {[names[i] for i in np.nonzero(bad)[0]]}"
        flags = np.array([out.terminated, out.base.truncated, out.success, out.route_ready, out.route_ready_streak,
                          out.route_regression, out.route_orientation_hit, out.route_index])
        assert np.array_equal(flags, g[prefix + "flags"][t]), where
        sc = np.array([out.route_q_error_norm, out.nearest_route_q_distance, out.base.position_error_norm, out.base.orientation_error_norm])
        assert np.abs(sc - g[prefix + "scalars"][t]).max() < 1e-12, where
        assert np.abs(obs - g[prefix + "obs"][t]).max() <= 1.2e-7, where

    env = ko.OracleRouteEnv(params, route)
    starts = g["seq_episode_start"]
    for e in range(len(starts) - 1):
        obs = env.reset(route_index=int(g["seq_reset_index"][e]), start_route_index=0, initial_q=g["seq_reset_q"][e],
                        initial_dq=g["seq_reset_dq"][e], initial_prev_action=g["seq_reset_pa"][e])
        assert np.abs(obs - g["seq_reset_obs"][e]).max() <= 1.2e-7
        for t in range(starts[e], starts[e + 1]):
            obs, out = env.step(g["seq_action"][t])
            check("seq_", t, obs, out)

    senv = ko.OracleRouteEnv(params, route, sequence_length=4, max_route_index=len(route) - 1)
    starts = g["adv_episode_start"]
    for e in range(len(starts) - 1):
        first = int(g["adv_reset_index"][e])
        obs = senv.reset(route_index=first, start_route_index=first - 1)
        assert np.abs(obs - g["adv_reset_obs"][e]).max() <= 1.2e-7
        for t in range(starts[e], starts[e + 1]):
            obs, out = senv.step(g["adv_action"][t])
            check("adv_", t, obs, out)
            assert senv.state.completed_waypoints == g["adv_completed"][t]


def test_route_sequential_probe_prefix():
    g = golden("trace_route.npz")
    params, _, _ = _route_setup()
    route = ko.OracleRoute(g["route_q"])
    prefix, flags, errs, steps = ko.route_sequential_probe(params, route, oracle_policy("route_prefix120"), start_index=1, end_index=12)
    assert flags.shape == (12,) and steps >= 12 and 0 <= prefix <= 12
    assert errs.shape == (12, 3)
