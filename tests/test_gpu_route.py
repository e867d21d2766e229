"""GPU parity of the dense holder-route wrappers (waypoint advance, route reward, 80-float obs, sequential probe) vs the oracle."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import kin_oracle as ko

from ._util import golden, oracle_params, oracle_policy

pytestmark = pytest.mark.gpu


def _setup():
    from rl_brain_trainer_b200 import config as kcfg
    from rl_brain_trainer_b200.route import RouteDataset

    g = golden("trace_route.npz")
    renv, seq = kcfg.to_route_env_config(kcfg.preset_dict("route_prefix120"), max_route_index=len(g["route_q"]) - 1)
    route = RouteDataset.from_q(g["route_q"])
    return g, renv, seq, route, oracle_params(renv.base_env_config, renv.reward_config)


def test_route_dataset_matches_reference():
    g, _, _, route, _ = _setup()
    assert np.abs(route.pose6 - g["route_pose6"]).max() < 1e-9
    assert np.abs(route.progress_m - g["route_progress"]).max() < 1e-9


def _compare_step(prefix, t, g, obs, reward, term, trunc, info, comps):
    where = f"{prefix} step {t}"
    flags = np.array([int(term[0]), int(trunc[0]), int(info["success"][0]), int(info["route_ready"][0]), int(info["route_ready_streak"][0]),
                      int(info["route_regression"][0]), int(info["route_orientation_hit"][0]), int(info["route_index"][0])])
    assert np.array_equal(flags, g[prefix + "flags"][t]), where
    sc = np.array([float(info["route_q_error_norm"][0]), float(info["nearest_route_q_distance"][0]), float(info["position_error_norm"][0]),
                   float(info["orientation_error_norm"][0])])
    assert np.abs(sc - g[prefix + "scalars"][t]).max() < 1e-5, where
    assert np.abs(obs[0].cpu().numpy() - g[prefix + "obs"][t]).max() < 5e-5, where
    ref = g[prefix + "components"][t]
    got = comps[:17, 0].cpu().numpy()
    assert np.abs(got - ref).max() < 3e-4 * max(1.0, np.abs(ref).max()), (where, got - ref)
    assert abs(float(reward[0]) - g[prefix + "reward"][t]) < 5e-4 * max(1.0, abs(g[prefix + "reward"][t])), where


def test_route_env_golden_sequential_chain():
    """RouteKinematicEnv with the evaluator's state override, 12 chained waypoints recorded from the reference."""
    from rl_brain_trainer_b200.route import BatchedRouteKinematicEnv

    g, renv, _, route, _ = _setup()
    env = BatchedRouteKinematicEnv(route, renv, 1, with_components=True)
    starts = g["seq_episode_start"]
    for e in range(len(starts) - 1):
        obs = env.reset(route_index=[int(g["seq_reset_index"][e])], start_route_index=[0], initial_q=g["seq_reset_q"][e],
                        initial_dq=g["seq_reset_dq"][e], initial_prev_action=g["seq_reset_pa"][e])
        assert np.abs(obs[0].cpu().numpy() - g["seq_reset_obs"][e]).max() < 5e-5
        for t in range(starts[e], starts[e + 1]):
            obs, reward, term, trunc, info = env.step(torch.as_tensor(g["seq_action"][t][None], dtype=torch.float32))
            _compare_step("seq_", t, g, obs, reward, term, trunc, info, info["reward_components"])


def test_route_sequence_env_golden_waypoint_advance():
    """RouteSequenceKinematicEnv: the target advances in place inside one episode (sequence_length 4)."""
    from dataclasses import replace

    from rl_brain_trainer_b200.route import BatchedRouteKinematicEnv

    g, renv, seq, route, _ = _setup()
    env = BatchedRouteKinematicEnv(route, renv, 1, sequence_config=replace(seq, enabled=True, sequence_length=4), with_components=True)
    starts = g["adv_episode_start"]
    for e in range(len(starts) - 1):
        first = int(g["adv_reset_index"][e])
        obs = env.reset(route_index=[first], start_route_index=[first - 1])
        assert np.abs(obs[0].cpu().numpy() - g["adv_reset_obs"][e]).max() < 5e-5
        for t in range(starts[e], starts[e + 1]):
            obs, reward, term, trunc, info = env.step(torch.as_tensor(g["adv_action"][t][None], dtype=torch.float32))
            _compare_step("adv_", t, g, obs, reward, term, trunc, info, info["reward_components"])
            assert int(info["route_completed_waypoints"][0]) == int(g["adv_completed"][t])


def test_route_batch_vs_oracle_random():
    """256 replicas at random waypoints with noisy proportional actions, 40 steps, sequence mode, vs the oracle."""
    from dataclasses import replace

    from rl_brain_trainer_b200.route import BatchedRouteKinematicEnv

    g, renv, seq, route, params = _setup()
    rng = np.random.default_rng(3)
    n, T = 256, 40
    env = BatchedRouteKinematicEnv(route, renv, n, sequence_config=replace(seq, enabled=True, sequence_length=3))
    oroute = ko.OracleRoute(g["route_q"])
    first = rng.integers(1, len(route) - 4, size=n)
    obs = env.reset(route_index=first, start_route_index=first - 1)
    oenvs = []
    for e in range(n):
        oe = ko.OracleRouteEnv(params, oroute, sequence_length=3, max_route_index=len(route) - 1)
        oe.reset(route_index=int(first[e]), start_route_index=int(first[e]) - 1)
        oenvs.append(oe)
    dl = np.array([s.delta_limit for s in renv.base_env_config.joint_specs]) * renv.base_env_config.action_delta_scale
    alive = np.ones(n, dtype=bool)
    mism = 0
    for t in range(T):
        q = env.state[0:7, :n].t().cpu().numpy().astype(float)
        idx = env.raux[5, :n].view(torch.int32).cpu().numpy() if t else first
        goal = g["route_q"][np.clip(idx, 0, len(route) - 1)]
        a = np.clip((goal - q) / dl * 0.7 + rng.normal(0, 0.03, (n, 7)), -1, 1).astype(np.float32)
        obs, reward, term, trunc, info = env.step(torch.as_tensor(a))
        for e in range(n):
            if not alive[e]:
                continue
            robs, ro = oenvs[e].step(a[e].astype(float))
            same = (int(term[e]) == ro.terminated and int(info["success"][e]) == ro.success and int(info["route_index"][e]) == ro.route_index
                    and int(info["route_ready"][e]) == ro.route_ready)
            if not same:           # threshold-adjacent flip: stop comparing this replica (its trajectory of targets diverges)
                mism += 1
                alive[e] = False
                continue
            assert np.abs(obs[e].cpu().numpy() - robs).max() < 1e-4
            assert abs(float(reward[e]) - ro.route_reward) < 1e-3 * max(1.0, abs(ro.route_reward))
            if ro.terminated or ro.base.truncated:
                alive[e] = False
    assert mism <= 0.03 * n


def test_route_sequential_probe_matches_oracle():
    from rl_brain_trainer_b200.policy import PolicyWeights
    from rl_brain_trainer_b200.route import evaluate_sequential_route, synthetic_route

    g, renv, _, route, params = _setup()
    pol = PolicyWeights.preset("route_prefix120", "cuda")
    res = evaluate_sequential_route(route, renv, pol, n_replicas=64, start_index=1, end_index=len(route) - 1, start_q_noise_std=0.0008, seed=1)
    prefix, flags, errs, steps = ko.route_sequential_probe(params, ko.OracleRoute(g["route_q"]), oracle_policy("route_prefix120"), start_index=1,
                                                           end_index=len(route) - 1)
    bits = res["success_bits"][0].cpu().numpy().view(np.uint32)
    got = np.array([(bits[k >> 5] >> (k & 31)) & 1 for k in range(len(flags))])
    assert np.mean(got != flags) <= 0.08, (got, flags)
    assert abs(res["replica0_longest_success_prefix"] - prefix) <= 2
    assert int(res["prefix_histogram"].sum()) == 64
    # a longer seeded synthetic route in the reference's dimensions runs end to end
    big = synthetic_route(483, seed=7)
    res = evaluate_sequential_route(big, renv, pol, n_replicas=256, start_index=1, end_index=170, start_q_noise_std=0.0008)
    assert int(res["env_steps"].item()) >= 170 * 256 and res["longest_success_prefix"].shape == (256,)


def test_route_probe_tensor_core_variant_tracks_the_strict_probe():
    """kin_route_probe_tc (80-input actor on tcgen05, TF32 operands) against the strict-fp32 probe on the same replicas: the waypoint
    success flags agree up to the O(1e-3) action differences of TF32 + tanh.approx, for full and partial tiles / several CTAs."""
    from rl_brain_trainer_b200.policy import PolicyWeights
    from rl_brain_trainer_b200.route import evaluate_sequential_route, synthetic_route

    _, renv, _, _, _ = _setup()
    pol = PolicyWeights.preset("route_prefix120", "cuda")
    route = synthetic_route(483, seed=7)
    for n in (1, 200, 20000):       # one partial tile; two tiles in one CTA; one CTA per SM with one tile each
        kw = dict(n_replicas=n, start_index=1, end_index=40, start_q_noise_std=0.0008, seed=3)
        a = evaluate_sequential_route(route, renv, pol, variant="fp32", **kw)
        b = evaluate_sequential_route(route, renv, pol, variant="tc", **kw)
        ba, bb = a["success_bits"].cpu().numpy().view(np.uint32), b["success_bits"].cpu().numpy().view(np.uint32)
        flips = sum(bin(int(x)).count("1") for x in (ba ^ bb).reshape(-1))
        assert flips <= 0.02 * n * 40 + 2, (n, flips)
        pa, pb = a["longest_success_prefix"].float(), b["longest_success_prefix"].float()
        assert abs(float(pa.mean()) - float(pb.mean())) <= 0.05 * 40 + 1
        sa, sb = int(a["env_steps"].item()), int(b["env_steps"].item())
        assert abs(sa - sb) <= 0.03 * sa
        assert int(b["prefix_histogram"].sum()) == n
    # results depend on the replica alone, not on the launch shape: replica 0 (no start noise) gives the same bits in every batch
    one = evaluate_sequential_route(route, renv, pol, variant="tc", n_replicas=1, start_index=1, end_index=40)
    many = evaluate_sequential_route(route, renv, pol, variant="tc", n_replicas=777, start_index=1, end_index=40, start_q_noise_std=0.0008, seed=9)
    assert torch.equal(one["success_bits"][0], many["success_bits"][0])


def test_sampled_route_reset_and_route_window():
    """reset() without an explicit waypoint draws from the batched device port of sample_route_reset; set_route_window narrows it."""
    from rl_brain_trainer_b200.route import BatchedRouteKinematicEnv, _reset_mode_ratios

    g, renv, _, route, _ = _setup()
    n = 16384
    env = BatchedRouteKinematicEnv(route, renv, n)
    env.set_route_window(max_route_index=30, min_route_index=4)
    obs = env.reset(seed=3)
    torch.cuda.synchronize()
    smp = env.last_reset
    cfg = env.config.reset_config
    assert obs.shape == (n, 80) and bool(torch.isfinite(obs).all())
    frac = torch.bincount(smp["reset_mode"], minlength=5).double().cpu().numpy() / n
    assert np.abs(frac - _reset_mode_ratios(cfg)).max() < 0.015
    ri, si, mode = smp["route_index"].long(), smp["start_route_index"].long(), smp["reset_mode"]
    assert int(ri.min()) >= 1 and int(ri.max()) <= 30
    plain = (mode == 1) | (mode == 4) | (mode == 0)
    assert int(ri[plain].min()) == 4 and int(ri[plain].max()) == 30           # uniform over the window, both ends reached
    assert bool((si[mode == 0] == 0).all()) and bool((si[mode != 0] == ri[mode != 0] - 1).all())
    # starts: the waypoint before the target (the target itself in recovery mode) + N(0, q_noise_std), clipped
    src = torch.where(mode == 4, ri, si)
    dq0 = smp["initial_q"] - env.table.q[src]
    assert abs(float(dq0.std()) - cfg.q_noise_std) < 0.1 * cfg.q_noise_std and float(dq0.abs().max()) < 6 * cfg.q_noise_std
    assert abs(float(smp["initial_dq"].std()) - cfg.dq_noise_std) < 0.1 * cfg.dq_noise_std
    assert float(smp["initial_prev_action"].abs().max()) <= 1.0
    info_idx = env.state  # the reset landed in the env state: q row equals the sampled start
    assert torch.allclose(info_idx[0:7, :n].t(), smp["initial_q"], atol=1e-6)
    # same seed -> same draws; the window can be moved between resets (RoutePrefixCurriculum promotion)
    env.reset(seed=3)
    assert torch.equal(env.last_reset["route_index"], smp["route_index"])
    env.set_route_window(max_route_index=8)
    env.reset()
    assert int(env.last_reset["route_index"].max()) <= 8 and int(env.last_reset["route_index"].min()) >= 1
    # explicit waypoints still bypass the sampler
    env.reset(route_index=[5])
    assert env.last_reset is None


def test_device_route_auto_reset_matches_the_sampler_distribution():
    """kin_route_reset_sampled (the route env's auto-reset, one launch): only the finished slots are touched; the target waypoints,
    the start-state noise and the reset-mode mix follow sample_route_reset; the sequence wrapper's window bookkeeping is applied."""
    import dataclasses

    from rl_brain_trainer_b200 import _lib
    from rl_brain_trainer_b200 import config as kcfg
    from rl_brain_trainer_b200.route import BatchedRouteKinematicEnv, _reset_index_ranges, _reset_mode_ratios

    g, renv, _, route, _ = _setup()
    n = 32768
    D = _lib.define
    for sequence in (False, True):
        seq = kcfg.RouteSequenceConfig(enabled=True, sequence_length=3) if sequence else None
        env = BatchedRouteKinematicEnv(route, renv, n, sequence_config=seq)
        env.set_route_window(max_route_index=30, min_route_index=4)
        env.reset(route_index=[7])
        before_state, before_obs = env.state.clone(), env.obs.clone()
        done = torch.zeros(n, dtype=torch.uint8, device="cuda")
        done[::2] = D("KIN_DONE_TERMINATED")
        done[1::4] = D("KIN_DONE_SUCCESS")            # a success flag alone is not "finished"
        env.reset_done(seed=11, counter=5, done=done)
        torch.cuda.synchronize()
        hit = torch.arange(n, device="cuda") % 2 == 0
        assert torch.equal(env.state[:, :n][:, ~hit], before_state[:, :n][:, ~hit]) and torch.equal(env.obs[~hit], before_obs[~hit])
        ri = (env.state[D("KIN_ROW_ROUTE"), :n].view(torch.int32) & 0xFFFF).long()[hit]
        last = (env.state[D("KIN_ROW_ROUTE2"), :n].view(torch.int32) & 0xFFFF).long()[hit]
        cfg = env.config.reset_config
        ranges = _reset_index_ranges(cfg, len(route) - 1)
        assert int(ri.min()) >= int(ranges[:, 0].min()) and int(ri.max()) <= 30
        if sequence:
            assert bool((last == torch.clamp(ri + 2, max=30)).all())
        else:
            assert bool((last == ri).all())
        # start state = waypoint(src) + N(0, q_noise_std): src is target - 1 (or 0 / the target itself for two of the modes)
        q = env.state[0:7, :n].t()[hit]
        tab = env.table.q
        d_prev = (q - tab[(ri - 1).clamp_min(0)]).norm(dim=1)
        d_zero = (q - tab[torch.zeros_like(ri)]).norm(dim=1)
        d_self = (q - tab[ri]).norm(dim=1)
        nearest = torch.stack([d_zero, d_prev, d_self]).min(dim=0).values
        bound = 6.0 * cfg.q_noise_std * 7 ** 0.5
        assert float(nearest.max()) < bound
        ratios = _reset_mode_ratios(cfg)
        frac_self = float(((d_self < bound) & (d_prev > bound)).float().mean())          # recovery mode starts AT the target
        assert abs(frac_self - ratios[4]) < 0.02
        # dq rows carry the dq noise, prev_action rows the clipped action noise
        dq = env.state[D("KIN_ROW_DQ"):D("KIN_ROW_DQ") + 7, :n].t()[hit]
        assert abs(float(dq.std()) - cfg.dq_noise_std) < 0.1 * cfg.dq_noise_std + 1e-9
        assert float(env.state[D("KIN_ROW_PREV_ACTION"):D("KIN_ROW_PREV_ACTION") + 7, :n].abs().max()) <= 1.0
        assert bool(torch.isfinite(env.obs).all())
        # a different counter gives different draws, the same one the same draws
        s1 = env.state.clone()
        env.reset_done(seed=11, counter=5, done=done)
        assert torch.equal(env.state, s1)
        env.reset_done(seed=11, counter=6, done=done)
        assert not torch.equal(env.state, s1)
        if not sequence:      # explicit resets: a negative waypoint is a masked (skipped) entry
            s2, o2 = env.state.clone(), env.obs.clone()
            mixed = torch.full((n,), -1, dtype=torch.int32, device="cuda")
            mixed[:3] = 9
            env.reset(route_index=mixed)
            assert torch.equal(env.state[:, 3:n], s2[:, 3:n]) and torch.equal(env.obs[3:], o2[3:])
            assert bool(((env.state[D("KIN_ROW_ROUTE"), :3].view(torch.int32) & 0xFFFF) == 9).all())
        # narrowing the window is picked up (the parameter block is rebuilt)
        env.set_route_window(max_route_index=8)
        env.reset_done(seed=11, counter=7, done=done)
        ri8 = (env.state[D("KIN_ROW_ROUTE"), :n].view(torch.int32) & 0xFFFF).long()[hit]
        assert int(ri8.max()) <= 8 and int(ri8.min()) >= 1


def _flags_of(term, trunc, info, index_key="route_index"):
    return np.array([int(term), int(trunc), int(info["success"]), int(info["route_ready"]), int(info["route_ready_streak"]),
                     int(info["route_regression"]), int(info["route_orientation_hit"]), int(info[index_key])])


def _flat80(obs):
    from rl_brain_trainer_b200.route import ROUTE_OBS_SLICES

    return np.concatenate([np.asarray(obs[k], dtype=np.float32).reshape(-1) for k in sorted(ROUTE_OBS_SLICES)])


def test_single_env_route_adapters_follow_the_reference_traces():
    """The 1-env drop-ins for RouteKinematicEnv / RouteSequenceKinematicEnv driven exactly as the reference's evaluator drives them
    (explicit reset, base_env.reset state override, private attribute writes, _augment_obs) against traces recorded from the live
    reference: dict observations, rewards, flags, info scalars."""
    from dataclasses import replace

    from rl_brain_trainer_b200.route import RouteKinematicEnv, RouteSequenceKinematicEnv

    g, renv, seq, route, _ = _setup()
    env = RouteKinematicEnv(route=route, config=renv, seed=0)
    assert set(env.observation_space.spaces) >= {"route_q_goal", "route_q_error", "route_tangent", "route_scalar", "q", "dq"}
    starts = g["seq_episode_start"]
    for e in range(len(starts) - 1):
        gi = int(g["seq_reset_index"][e])
        obs, info = env.reset(options={"route_index": gi, "start_route_index": 0, "policy_mode": "approach"})
        assert info["route_index"] == gi and info["route_reset_mode"] == "explicit" and obs["route_q_goal"].shape == (7,)
        # eval_route_curriculum.py:73-87: override the base state for sequential chaining
        obs, info = env.base_env.reset(options={"initial_q": g["seq_reset_q"][e], "initial_dq": g["seq_reset_dq"][e],
                                                "initial_prev_action": g["seq_reset_pa"][e], "goal_q": route.q_goal[gi], "policy_mode": "approach"})
        env._route_index, env._start_route_index, env._ready_streak, env._prev_info = gi, 0, 0, dict(info)
        obs = env._augment_obs(obs)
        assert np.abs(_flat80(obs) - g["seq_reset_obs"][e]).max() < 5e-5
        for t in range(starts[e], starts[e + 1]):
            obs, reward, term, trunc, info = env.step(g["seq_action"][t])
            assert np.array_equal(_flags_of(term, trunc, info), g["seq_flags"][t]), t
            sc = np.array([info["route_q_error_norm"], info["nearest_route_q_distance"], info["position_error_norm"], info["orientation_error_norm"]])
            assert np.abs(sc - g["seq_scalars"][t]).max() < 1e-5 and np.abs(_flat80(obs) - g["seq_obs"][t]).max() < 5e-5, t
            assert abs(reward - g["seq_reward"][t]) < 5e-4 * max(1.0, abs(g["seq_reward"][t])) and isinstance(reward, float)
            assert np.abs(info["q"] - g["seq_q"][t]).max() < 5e-6 and np.abs(env.base_env._q - g["seq_q"][t]).max() < 5e-6
            assert env._ready_streak == int(g["seq_flags"][t][4]) and env._prev_info is not None
    with pytest.raises(ValueError):
        env.step(np.zeros(6))
    # sampled resets draw from the numpy stream like the reference; the window can be narrowed
    env.set_route_window(max_route_index=6)
    _, info = env.reset(seed=3)
    assert 1 <= info["route_index"] <= 6 and info["route_reset_mode"] in ("prefix_start", "random_prefix", "segment", "replay", "recovery")
    # the sequence wrapper: in-episode waypoint advance
    senv = RouteSequenceKinematicEnv(route=route, config=renv, sequence_config=replace(seq, enabled=True, sequence_length=4))
    starts = g["adv_episode_start"]
    for e in range(len(starts) - 1):
        first = int(g["adv_reset_index"][e])
        obs, info = senv.reset(options={"route_index": first, "start_route_index": first - 1})
        assert np.abs(_flat80(obs) - g["adv_reset_obs"][e]).max() < 5e-5 and info["route_last_index"] == min(first + 3, len(route) - 1)
        for t in range(starts[e], starts[e + 1]):
            obs, reward, term, trunc, info = senv.step(g["adv_action"][t])
            assert np.array_equal(_flags_of(term, trunc, info), g["adv_flags"][t]), t
            assert np.abs(_flat80(obs) - g["adv_obs"][t]).max() < 5e-5 and abs(reward - g["adv_reward"][t]) < 5e-4 * max(1.0, abs(g["adv_reward"][t]))
            assert info["route_completed_waypoints"] == int(g["adv_completed"][t]) == senv._completed_waypoints
            assert senv._current_route_index == int(g["adv_flags"][t][7])


def test_sequential_route_eval_rows_and_summary_match_the_reference():
    """evaluate_sequential_route with per-waypoint rows against the live reference's run of the same chain (tests/golden/route_eval.json,
    made by gen_golden_route_eval.py): _roll_one rows, _summarize_rows, _chunk_metrics, _failure_reason."""
    import json

    from rl_brain_trainer_b200.policy import PolicyWeights
    from rl_brain_trainer_b200.route import evaluate_sequential_route, route_failure_reason

    from ._util import GOLD

    ref = json.loads((GOLD / "route_eval.json").read_text())
    g, renv, _, route, _ = _setup()
    pol = PolicyWeights.preset("route_prefix120", "cuda")
    res = evaluate_sequential_route(route, renv, pol, n_replicas=3, start_index=1, end_index=len(route) - 1, start_q_noise_std=0.0008, seed=1,
                                    detail_replicas=2)
    rows, rrows = res["rows"][0], ref["rows"]
    assert len(res["rows"]) == 2 and len(rows) == len(rrows) == len(route) - 1
    # the chain is closed-loop: fp32 vs fp64 rounding can move a borderline waypoint, after which the two chains differ legitimately
    same = 0
    for a, b in zip(rows, rrows):
        if a["success"] != b["success"] or a["steps"] != b["steps"]:
            break
        same += 1
        assert a["route_index"] == b["route_index"] and a["route_ready_hit"] == b["route_ready_hit"] and a["route_ready_dwell"] == b["route_ready_dwell"]
        assert a["first_ready_step"] == b["first_ready_step"] and a["max_ready_streak"] == b["max_ready_streak"]
        for k in ("final_position_error", "final_orientation_error", "final_q_error", "min_position_error", "min_orientation_error", "min_q_error",
                  "final_action_magnitude", "final_dq_norm"):
            assert abs(a[k] - b[k]) < 2e-4 * max(1.0, abs(b[k])) + 2e-5, (a["route_index"], k, a[k], b[k])
        if not b["success"]:
            assert route_failure_reason(a) == ref["failure_reasons"][a["route_index"] - 1]
    assert same >= ref["summary"]["longest_success_prefix"] + 1          # at least through the first failure
    s, rs = res["summary"], ref["summary"]
    assert set(rs) <= set(s)
    assert s["longest_success_prefix"] == rs["longest_success_prefix"] == res["replica0_longest_success_prefix"]
    assert s["first_failure_index"] == rs["first_failure_index"] and s["first_failure_reason"] == rs["first_failure_reason"]
    assert abs(s["cumulative_successful_route_distance_m"] - rs["cumulative_successful_route_distance_m"]) < 1e-9
    assert abs(s["success_rate"] - rs["success_rate"]) <= 3 / len(rows) and abs(s["route_ready_hit_rate"] - rs["route_ready_hit_rate"]) <= 3 / len(rows)
    assert res["failure_report"]["first_failure"]["route_index"] == ref["failure_report"]["first_failure"]["route_index"]
    assert set(res["chunk_metrics"]) == set(ref["chunk_metrics"])
    for name, c in ref["chunk_metrics"].items():
        assert res["chunk_metrics"][name]["target_count"] == c["target_count"]
        assert abs(res["chunk_metrics"][name]["success_rate"] - c["success_rate"]) <= 3 / c["target_count"]
    with pytest.raises(ValueError):
        evaluate_sequential_route(route, renv, pol, n_replicas=2, variant="tc", detail_replicas=1)


def test_pruned_nearest_waypoint_scan_is_the_full_scan():
    """The route reward's nearest-waypoint distance (route_env.py:135 scans the whole route): the step kernel's pruned outward scan
    (KinRouteTable.nearest_lb, triangle inequality) returns the full scan's value BIT FOR BIT -- on-route replicas, replicas pushed
    far off the route (the pruning table runs out and the kernel finishes with the plain scan) and every route index."""
    from rl_brain_trainer_b200 import config as kcfg
    from rl_brain_trainer_b200.route import BatchedRouteKinematicEnv, nearest_scan_bounds, synthetic_route

    route = synthetic_route(483, seed=7)
    renv, _ = kcfg.to_route_env_config(kcfg.preset_dict("route_prefix120"), max_route_index=len(route) - 1)
    n = 8192
    lb = nearest_scan_bounds(route.q_goal)
    d = np.linalg.norm(route.q_goal[:, None] - route.q_goal[None], axis=2)
    for k in (0, 1, 5, 63):       # the table is a lower bound of the true minimum, and tight to fp32 rounding
        true = np.where(np.abs(np.arange(483)[:, None] - np.arange(483)[None]) >= k, d, np.inf).min(1)
        assert (lb[:, k] <= true).all() and (lb[:, k] >= true * (1 - 1e-6)).all()
    g = torch.Generator(device="cuda").manual_seed(3)
    idx = torch.randint(1, len(route), (n,), generator=g, device="cuda").cpu().numpy()
    out = []
    for pruned in (True, False):
        env = BatchedRouteKinematicEnv(route, renv, n)
        if not pruned:
            env.table.c.nearest_lb, env.table.c.nearest_lb_k = None, 0
        gq = torch.Generator(device="cuda").manual_seed(5)
        q0 = torch.as_tensor(route.q_goal[np.maximum(idx - 1, 0)], dtype=torch.float32, device="cuda")
        spread = torch.cat([torch.full((n // 2,), 0.01), torch.full((n // 4,), 0.3), torch.full((n - n // 2 - n // 4,), 2.0)]).to("cuda")[:, None]
        q0 = q0 + spread * torch.randn((n, 7), generator=gq, device="cuda")
        env.reset(route_index=idx, start_route_index=np.maximum(idx - 1, 0), initial_q=q0)
        rows = []
        for _ in range(3):
            a = torch.rand((n, 7), generator=gq, device="cuda") * 2 - 1
            _, reward, _, _, info = env.step(a)
            rows.append((info["nearest_route_q_distance"].clone(), reward.clone()))
        out.append(rows)
    for (na, ra), (nb, rb) in zip(*out):
        assert torch.equal(na, nb) and torch.equal(ra, rb)
    assert float(out[0][0][0].max()) > 0.5 and float(out[0][0][0].min()) < 0.05     # both regimes were exercised
