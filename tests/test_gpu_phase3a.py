"""Phase-3A Gazebo bridge helpers (rl_brain_trainer_b200/phase3a.py) against vectors produced by the live reference
(tests/golden/gen_golden_phase3a.py): FK, the observation builder for externally measured joint states, pose errors."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def test_phase3a_helpers_match_the_reference():
    from rl_brain_trainer_b200 import phase3a
    from rl_brain_trainer_b200.env import OBS_KEYS

    g = np.load(GOLD / "phase3a.npz")
    ctx = phase3a.Phase3AKinematics()
    specs = phase3a.default_joint_specs()
    keys = sorted(OBS_KEYS)                       # SB3 flattens the Dict observation in alphabetical key order
    n = g["q"].shape[0]
    for i in range(n):
        pose = ctx.compute_ee_pose6(g["q"][i])
        d = pose - g["current_pose6"][i]
        d[3:] = (d[3:] + np.pi) % (2 * np.pi) - np.pi
        assert np.abs(d[:3]).max() < 1e-5 and np.abs(d[3:]).max() < 1e-5          # north_star: 1e-5 m / 1e-5 rad in fp32
        wp = bool(g["use_wp"][i])
        obs = ctx.build_observation(q=g["q"][i], dq=g["dq"][i], prev_action=g["prev_action"][i], current_pose6=g["current_pose6"][i],
                                    goal_pose6=g["goal_pose6"][i], joint_specs=specs, episode_progress=float(g["episode_progress"][i]),
                                    dwell_progress=float(g["dwell_progress"][i]), mode_index=int(g["mode_index"][i]),
                                    current_waypoint_pose6=g["wp_pose6"][i] if wp else None, next_waypoint_pose6=g["next_wp_pose6"][i] if wp else None)
        assert set(obs) == set(OBS_KEYS) and all(v.dtype == np.float32 for v in obs.values())
        flat = np.concatenate([obs[k].reshape(-1) for k in keys])
        assert np.abs(flat - g["obs56"][i]).max() < 5e-6, (i, np.abs(flat - g["obs56"][i]).argmax())
        pe, oe = ctx.pose_error_components(g["current_pose6"][i], g["goal_pose6"][i])
        assert np.abs(pe - g["pos_err"][i]).max() < 1e-12 and np.abs(oe - g["ori_err"][i]).max() < 1e-12
    # the module-level functions (what the bridge imports) share one default context
    assert np.allclose(phase3a.compute_ee_pose6(g["q"][0]), ctx.compute_ee_pose6(g["q"][0]))
    with pytest.raises(ValueError):
        ctx.compute_ee_pose6(np.zeros(6))
