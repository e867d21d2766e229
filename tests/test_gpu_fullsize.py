"""Size-independent properties at BASELINE.json's full sizes (the oracle only finishes small cases in seconds):
determinism, permutation equivariance, shard invariance, ragged tails -- plus a sampled oracle check of the big batch."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from ._util import env_config

pytestmark = pytest.mark.gpu

N_FULL = 65536


def _rollout(variant):
    from rl_brain_trainer_b200.policy import PolicyWeights
    from rl_brain_trainer_b200.rollout import ApproachFinisherRollout

    return ApproachFinisherRollout(env_config("approach_dynamic_scale_big"), PolicyWeights.preset("approach_stage8_11", "cuda"),
                                   env_config("finisher_noop_ft"), PolicyWeights.preset("finisher", "cuda"), variant=variant)


def _suite(n):
    from rl_brain_trainer_b200.samplers import build_curriculum_local_eval_suite

    return build_curriculum_local_eval_suite(env_config("approach_dynamic_scale_big"), seed=700001 + 5 * 1009, stage_index=5, n_episodes=n)


def _sub(suite, idx):
    from rl_brain_trainer_b200.samplers import EvalSuite

    return EvalSuite(initial_q=suite.initial_q[idx], goal_q=suite.goal_q[idx])


@pytest.mark.parametrize("variant", [0, 1])
def test_stage5_full_size_properties(variant):
    """Config 2 (65 536 Stage-5 episodes): every episode's result is a function of that episode alone."""
    ro = _rollout(variant)
    suite = _suite(N_FULL)
    a = ro.evaluate_suite(suite)
    raw = a.raw[:, :N_FULL].clone()
    steps = int(a.env_steps.item())
    # deterministic
    b = ro.evaluate_suite(suite)
    assert torch.equal(raw, b.raw[:, :N_FULL]) and steps == int(b.env_steps.item())
    # permutation equivariance (an episode may sit in any thread / GEMM row / CTA)
    perm = np.random.default_rng(3).permutation(N_FULL)
    c = ro.evaluate_suite(_sub(suite, perm))
    assert torch.equal(c.raw[:, :N_FULL], raw[:, torch.as_tensor(perm, device="cuda")])
    # shard invariance (the multi-GPU split) and ragged tails (partial tiles / warps)
    for lo, hi in ((0, N_FULL // 2), (N_FULL // 2, N_FULL), (1000, 1000 + 4096 + 37), (N_FULL - 129, N_FULL)):
        d = ro.evaluate_suite(_sub(suite, slice(lo, hi)))
        assert torch.equal(d.raw[:, : hi - lo], raw[:, lo:hi]), (lo, hi)
    # whole-suite statistics: the reference's Stage-5 rate is ~0.93-0.97 on its 64-episode draws; 0.968 on the full suite
    res = a.to_numpy()
    assert 0.955 < res["success"].mean() < 0.98
    assert steps == int(res["approach_steps"].sum() + res["finisher_steps"].sum()) and steps <= N_FULL * 164


def test_stage5_full_size_variants_agree_and_sampled_oracle_check():
    """The tensor-core rollout (TF32 MLP) against the strict-fp32 one on all 65 536 episodes, and 512 episodes sampled from the
    whole suite against the CPU oracle."""
    from oracle import kin_oracle as ko

    from ._util import oracle_params, oracle_policy

    suite = _suite(N_FULL)
    strict, tc = _rollout(0).evaluate_suite(suite).to_numpy(), _rollout(1).evaluate_suite(suite).to_numpy()
    flips = int(np.sum(strict["success"] != tc["success"]))
    assert flips < 0.001 * N_FULL, flips
    assert abs(strict["success"].mean() - tc["success"].mean()) < 5e-4
    assert abs(strict["final_position_error"].mean() - tc["final_position_error"].mean()) < 2e-5
    idx = np.sort(np.random.default_rng(5).choice(N_FULL, 512, replace=False))
    ref, _ = ko.eval_approach_finisher(oracle_params(env_config("approach_dynamic_scale_big")), oracle_params(env_config("finisher_noop_ft")),
                                       oracle_policy("approach_stage8_11"), oracle_policy("finisher"),
                                       initial_q=suite.initial_q[idx].astype(np.float32).astype(float),
                                       goal_q=suite.goal_q[idx].astype(np.float32).astype(float), n_threads=8)
    assert int(np.sum(strict["success"][idx].astype(int) != ref["success"])) <= 2
    same = strict["success"][idx].astype(int) == ref["success"]
    assert np.quantile(np.abs(strict["final_position_error"][idx][same] - ref["final_position_error"][same]), 0.99) < 2e-5
    assert np.array_equal(strict["approach_steps"][idx], ref["approach_steps"])


def test_step_kernel_full_size_matches_small_batches():
    """K1 at an HBM-resident size (2 097 152 envs): any sampled env steps exactly as it does in a small batch."""
    from rl_brain_trainer_b200.env import BatchedArmKinematicEnv

    cfg = env_config("approach_dynamic_scale_big")
    n, k = 2_097_152, 4096
    g = torch.Generator(device="cuda").manual_seed(9)
    iq = (torch.rand((n, 7), device="cuda", generator=g) - 0.5) * 0.6
    gq = (torch.rand((n, 7), device="cuda", generator=g) - 0.5) * 0.6
    acts = [torch.randn((n, 7), device="cuda", generator=g) * 0.8 for _ in range(3)]
    big = BatchedArmKinematicEnv(cfg, n, "cuda", with_aux=False)
    big.reset(options={"initial_q": iq, "goal_q": gq})
    idx = torch.randperm(n, device="cuda", generator=g)[:k]
    idx[:3] = torch.tensor([0, n - 1, n - 33], device="cuda")          # first env, last env, last full warp
    small = BatchedArmKinematicEnv(cfg, k, "cuda", with_aux=False)
    small.reset(options={"initial_q": iq[idx], "goal_q": gq[idx]})
    for a in acts:
        ob, rb, tb, ub, _ = big.step(a)
        os_, rs, ts, us, _ = small.step(a[idx].contiguous())
        assert torch.equal(ob[idx], os_) and torch.equal(rb[idx], rs) and torch.equal(tb[idx], ts) and torch.equal(ub[idx], us)
    assert torch.equal(big.state[:, idx], small.state[:, :k])
