"""Host-side helpers of the Phase-3A shim (no GPU): joint-limit arithmetic and the bridge's command / delta-scale rules."""
from pathlib import Path

import numpy as np

GOLD = Path(__file__).resolve().parent / "golden"


def test_joint_limit_helpers_and_command_rules():
    from rl_brain_trainer_b200 import phase3a

    g = np.load(GOLD / "phase3a.npz")
    specs = phase3a.default_joint_specs()
    assert np.array_equal(phase3a.delta_limits(specs), g["delta_limits"])
    for i in range(g["q"].shape[0]):
        assert np.array_equal(phase3a.clip_joint_configuration(g["q"][i], specs), g["clipped_q"][i])
    q = g["clipped_q"][0]
    cmd = phase3a.action_to_command_q(q=q, action=np.full(7, 2.0), joint_specs=specs, action_delta_scale=0.5)
    assert np.allclose(cmd, phase3a.clip_joint_configuration(q + 0.5 * phase3a.delta_limits(specs), specs))
    cfg = dict(action_delta_scale=1.0, dynamic_action_delta_scale_enabled=True, dynamic_action_delta_scale_near_pos_threshold_m=0.02,
               dynamic_action_delta_scale_far_pos_threshold_m=0.10, dynamic_action_delta_scale_near_multiplier=0.04,
               dynamic_action_delta_scale_far_multiplier=0.11)
    # the reference's dynamic-limit endpoints (SURVEY 8c known answers 0.11 / 0.04) and the midpoint
    assert phase3a.effective_action_delta_scale(cfg, 0.5) == 0.11 and phase3a.effective_action_delta_scale(cfg, 0.001) == 0.04
    assert abs(phase3a.effective_action_delta_scale(cfg, 0.06) - 0.075) < 1e-12
    assert phase3a.effective_action_delta_scale({**cfg, "dynamic_action_delta_scale_enabled": False}, 0.5) == 1.0
