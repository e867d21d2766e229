#!/usr/bin/env python
"""Golden draws of the reference's route reset sampler and prefix-curriculum callback logic (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden_route_reset.py

Runs route/route_reset_samplers.py::sample_route_reset (imported from /root/reference) on the 40-waypoint synthetic route of
trace_route.npz with three configs, 200 draws each.  Output: tests/golden/route_reset.npz
"""
from __future__ import annotations

import json
import sys
import tempfile
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import gen_golden as gg  # noqa: E402

from hrl_trainer.kinematic_phase1.route.route_reset_samplers import sample_route_reset  # noqa: E402

CONFIGS = {
    "mixed": dict(max_route_index=30, segment_start_index=5, segment_end_index=20, replay_start_index=2, replay_end_index=12, replay_reset_ratio=0.1),
    "recovery": dict(mode="recovery_reset", min_route_index=3, max_route_index=39, q_noise_std=0.01),
    "prefix_nonoise": dict(mode="prefix_start_reset", max_route_index=25, q_noise_std=0.0, dq_noise_std=0.0, prev_action_noise_std=0.0),
}


def main() -> None:
    cfgs = gg.merged_configs()
    base_cfg = gg.to_env_config(cfgs["route_prefix120"])
    route_q = gg.synthetic_route()
    with tempfile.TemporaryDirectory() as td:
        path = Path(td) / "route.json"
        path.write_text(json.dumps({"route_q": route_q.tolist()}))
        route = gg.load_route_dataset(path)
    out = {"route_q": route_q}
    modes = ["prefix_start", "random_prefix", "segment", "replay", "recovery"]
    for name, kw in CONFIGS.items():
        cfg = gg.RouteResetSamplerConfig(**kw)
        rng = np.random.default_rng(123)
        rows = [sample_route_reset(rng=rng, route=route, joint_specs=base_cfg.joint_specs, config=cfg) for _ in range(200)]
        out[f"{name}_initial_q"] = np.array([r.initial_q for r in rows])
        out[f"{name}_initial_dq"] = np.array([r.initial_dq for r in rows])
        out[f"{name}_initial_prev_action"] = np.array([r.initial_prev_action for r in rows])
        out[f"{name}_goal_q"] = np.array([r.goal_q for r in rows])
        out[f"{name}_route_index"] = np.array([r.route_index for r in rows])
        out[f"{name}_start_route_index"] = np.array([r.start_route_index for r in rows])
        out[f"{name}_mode"] = np.array([modes.index(r.reset_mode) for r in rows])
        out[f"{name}_config"] = np.array(json.dumps(kw))
    np.savez_compressed(gg.GOLD / "route_reset.npz", **out)
    print("route_reset.npz written:", {k: int(np.bincount(out[f'{k}_mode'], minlength=5).max()) for k in CONFIGS})


if __name__ == "__main__":
    main()
