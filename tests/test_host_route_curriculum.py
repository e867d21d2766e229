"""CPU tests: host port of the route reset sampler vs draws of the live reference; prefix-curriculum promotion logic."""

from __future__ import annotations

import json

import numpy as np
import pytest

from rl_brain_trainer_b200 import config as kcfg
from rl_brain_trainer_b200.route import ROUTE_RESET_MODES, RouteCurriculumStage, RouteDataset, RoutePrefixCurriculum, sample_route_reset

from ._util import env_config, golden


def test_route_reset_sampler_matches_reference_stream():
    g = golden("route_reset.npz")
    route = RouteDataset.from_q(g["route_q"])
    specs = env_config("approach_default").joint_specs
    for name in ("mixed", "recovery", "prefix_nonoise"):
        cfg = kcfg.RouteResetSamplerConfig(**json.loads(str(g[f"{name}_config"])))
        rng = np.random.default_rng(123)
        for i in range(200):
            s = sample_route_reset(rng, route, specs, cfg)
            assert s["route_index"] == g[f"{name}_route_index"][i] and s["start_route_index"] == g[f"{name}_start_route_index"][i], (name, i)
            assert ROUTE_RESET_MODES.index(s["reset_mode"]) == g[f"{name}_mode"][i]
            for k in ("initial_q", "initial_dq", "initial_prev_action", "goal_q"):
                assert np.array_equal(s[k], g[f"{name}_{k}"][i]), (name, i, k)
    assert len(set(g["mixed_mode"].tolist())) == 5            # every branch was exercised


def _curriculum(**kw):
    stages = [RouteCurriculumStage("p20", 20), RouteCurriculumStage("p60", 60), RouteCurriculumStage("p120", 120)]
    args = dict(promotion_success_rate=0.8, promotion_route_ready_hit_rate=0.7, promotion_orientation_hit_rate=0.6, promotion_max_regression_rate=0.2,
                window_episodes=10, min_episodes_per_stage=12)
    args.update(kw)
    return RoutePrefixCurriculum(stages, **args)


def test_prefix_curriculum_promotion_rules():
    c = _curriculum()
    assert c.prefix_end_index == 20
    ok = lambda n: (np.ones(n), np.ones(n), np.ones(n), np.zeros(n))  # noqa: E731
    assert not c.record(*ok(11))                      # window full (10) but fewer than min_episodes_per_stage (12)
    assert c.record(*ok(1), total_timesteps=999)      # 12th episode promotes
    assert c.prefix_end_index == 60 and c.stage_episode_count == 0 and c.metrics()["recent_success_rate"] == 0.0
    assert c.history[0]["from_prefix_end_index"] == 20 and c.history[0]["to_prefix_end_index"] == 60 and c.history[0]["total_timesteps"] == 999
    # one failing criterion blocks the promotion: regression rate above the bound
    assert not c.record(np.ones(20), np.ones(20), np.ones(20), np.r_[np.ones(3), np.zeros(17)][::-1].copy() * 0 + np.tile([1, 0, 0], 7)[:20])
    assert c.prefix_end_index == 60
    # ... until the window holds good episodes only
    assert c.record(*ok(10)) and c.prefix_end_index == 120
    # last stage: nothing to promote to
    assert not c.record(*ok(40)) and c.summary()["stage_name"] == "p120" and len(c.history) == 2
    # orientation-hit criterion
    c2 = _curriculum(min_episodes_per_stage=1)
    assert not c2.record(np.ones(10), np.ones(10), np.r_[np.ones(5), np.zeros(5)], np.zeros(10))
    with pytest.raises(ValueError):
        RoutePrefixCurriculum([], promotion_success_rate=1, promotion_route_ready_hit_rate=1, promotion_orientation_hit_rate=1,
                              promotion_max_regression_rate=0, window_episodes=1)


def test_route_eval_summaries_reproduce_the_reference_on_its_own_rows():
    """summarize_route_rows / route_chunk_metrics / route_failure_reason (eval_route_curriculum.py:127-186) fed with the per-waypoint
    rows the live reference produced (tests/golden/route_eval.json) must give the reference's own summary, chunk table and reasons."""
    import json
    from pathlib import Path

    import numpy as np

    from rl_brain_trainer_b200.route import RouteDataset, route_chunk_metrics, route_failure_reason, summarize_route_rows

    gold = Path(__file__).resolve().parent / "golden"
    ref = json.loads((gold / "route_eval.json").read_text())
    route = RouteDataset.from_q(np.load(gold / "trace_route.npz")["route_q"])
    got = summarize_route_rows(ref["rows"], route)
    for k, v in ref["summary"].items():
        if isinstance(v, float):
            assert abs(got[k] - v) < 1e-9, k
        else:
            assert got[k] == v, k
    assert route_chunk_metrics(ref["rows"]) == ref["chunk_metrics"]
    assert [None if r["success"] else route_failure_reason(r) for r in ref["rows"]] == ref["failure_reasons"]
    assert summarize_route_rows([], route) == {"target_count": 0}
