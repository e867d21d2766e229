"""Host-side samplers / suite builders / workspace maps consume numpy's PCG64 stream like the reference.

Golden draws: tests/golden/samplers.npz, randomstart_maps.npz, eval_randomstart.npz (recorded from the live reference).
"""

from __future__ import annotations

import numpy as np

from oracle import kin_oracle as ko
from rl_brain_trainer_b200 import samplers, workspace

from ._util import env_config, golden


def _draw_resets(cfg, mode, stage, seed, n):
    rng = np.random.default_rng(seed)
    samplers.sample_reset(rng, cfg, mode, stage, fk=ko.fk_pose6)  # the reference's gen loop resets once with the seed first
    return [samplers.sample_reset(rng, cfg, mode, stage, fk=ko.fk_pose6) for _ in range(n)]


def test_approach_stage_mix_stream():
    g = golden("samplers.npz")
    cfg = env_config("approach_dynamic_scale_big")
    for stage in (0, 5, 11):
        draws = _draw_resets(cfg, "approach", stage, 1000 + stage, 64)
        assert np.array_equal(np.array([d.initial_q for d in draws]), g[f"approach_s{stage}_q"])
        assert np.array_equal(np.array([d.goal_q for d in draws]), g[f"approach_s{stage}_goal_q"])


def test_randomstart_pair_stream():
    g = golden("samplers.npz")
    cfg = env_config("randomstart_overnight")
    for stage in (8, 11):
        draws = _draw_resets(cfg, "approach", stage, 1000 + stage, 64)
        assert np.array_equal(np.array([d.initial_q for d in draws]), g[f"randomstart_s{stage}_q"])
        assert np.array_equal(np.array([d.goal_q for d in draws]), g[f"randomstart_s{stage}_goal_q"])
        assert np.array_equal(np.array([d.initial_dq for d in draws]), g[f"randomstart_s{stage}_dq"])
        assert np.array_equal(np.array([d.initial_prev_action for d in draws]), g[f"randomstart_s{stage}_prev_action"])
        # goal pose = FK(goal_q) on the oracle
        assert np.abs(ko.fk_pose6(g[f"randomstart_s{stage}_goal_q"]) - g[f"randomstart_s{stage}_goal_pose6"]).max() < 1e-12


def test_dock_reset_stream_with_close_bucket():
    g = golden("samplers.npz")
    cfg = env_config("finisher_noop_ft")  # close_bucket_probability 0.15 exercises the FK rejection loop
    draws = _draw_resets(cfg, "dock", 0, 77, 64)
    assert np.abs(np.array([d.initial_q for d in draws]) - g["dock_q"]).max() < 1e-15
    assert np.array_equal(np.array([d.goal_q for d in draws]), g["dock_goal_q"])


def test_curriculum_local_suite_vectorised_stream():
    g = golden("samplers.npz")
    cfg = env_config("approach_dynamic_scale_big")
    suite = samplers.build_curriculum_local_eval_suite(cfg, seed=700001 + 5 * 1009, stage_index=5, n_episodes=256)
    assert np.array_equal(suite.initial_q, g["suite5_initial_q"])
    assert np.array_equal(suite.goal_q, g["suite5_goal_q"])
    assert np.abs(ko.fk_pose6(suite.goal_q) - g["suite5_goal_pose6"]).max() < 1e-12
    # a stage without start noise takes the goal-only branch
    s0 = samplers.build_curriculum_local_eval_suite(cfg, seed=700001, stage_index=0, n_episodes=16)
    ev = golden("eval_stages.npz")
    assert np.array_equal(s0.initial_q, ev["s0_initial_q"]) and np.array_equal(s0.goal_q, ev["s0_goal_q"])


def test_workspace_maps_and_pairs():
    g = golden("randomstart_maps.npz")
    cfg = env_config("randomstart_overnight")
    seed = 940001
    targets = workspace.generate_workspace_target_map(cfg, seed=seed + 1, stage_samples_per_stage=96, random_samples=384)
    starts = workspace.generate_workspace_start_state_map(cfg, seed=seed + 2, stage_samples_per_stage=48, random_samples=384)
    assert np.array_equal(targets.q, g["target_q"]) and np.array_equal(targets.stage, g["target_stage"])
    assert np.array_equal(starts.q, g["start_q"]) and np.array_equal(starts.dq, g["start_dq"])
    assert np.array_equal(starts.prev_action, g["start_prev_action"])
    names = list(g["start_source_names"])
    assert [workspace.START_SOURCES[i] for i in starts.source] == [names[i] for i in g["start_source"]]
    pairs = workspace.build_pair_table(starts, targets, seed=seed + 3, pair_count=2048)
    assert np.array_equal(pairs.start, g["pair_start"]) and np.array_equal(pairs.target, g["pair_target"])
    assert np.array_equal(pairs.klass, g["pair_class"])
    assert np.abs(pairs.q_l2 - g["pair_q_l2"]).max() < 1e-12


def test_randomstart_splits_match_reference_selection():
    ev = golden("eval_randomstart.npz")
    cfg = env_config("randomstart_overnight")
    suites = workspace.build_randomstart_eval(cfg, seed=940001)
    for split in ("known", "frontier", "stress"):
        s = suites[split]
        assert np.array_equal(s.initial_q, ev[f"{split}_initial_q"]) and np.array_equal(s.goal_q, ev[f"{split}_goal_q"])
        assert np.array_equal(s.initial_dq, ev[f"{split}_initial_dq"])
        assert np.array_equal(s.initial_prev_action, ev[f"{split}_initial_prev_action"])


def test_dock_eval_suite_matches_reference_stream():
    """build_dock_eval_suite (fixed_eval_suite.py:108-134) on the finisher config: same PCG64 stream as the live reference, including
    the close-bucket rejection sampler's FK calls (tests/golden/gen_golden_dock_suite.py)."""
    from pathlib import Path

    from rl_brain_trainer_b200 import config as kcfg, kinematics
    from rl_brain_trainer_b200.samplers import build_dock_eval_suite

    g = np.load(Path(__file__).resolve().parent / "golden" / "dock_suite.npz")
    cfg = kcfg.load_preset("finisher_noop_ft")
    suite = build_dock_eval_suite(cfg, seed=700001, n_episodes=24, fk=kinematics.fk_pose6_folded)
    assert len(suite) == 24
    assert np.abs(suite.initial_q - g["initial_q"]).max() < 1e-12 and np.abs(suite.goal_q - g["goal_q"]).max() < 1e-12
