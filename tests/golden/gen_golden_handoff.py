#!/usr/bin/env python
"""Golden fixture for the handoff-state buffer builder, produced by RUNNING THE LIVE REFERENCE (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden_handoff.py

Runs the loop of kinematic_phase1/training/build_finisher_handoff_state_buffer.py:73-125 (its own helpers, imported from
/root/reference) for 48 Stage-10 episodes of the bundled approach checkpoint and records, per episode, what each of the three
handoff modes would store.  Output: tests/golden/handoff_states.npz
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import gen_golden as gg  # noqa: E402  (puts the reference on sys.path, loads checkpoints / configs)

from hrl_trainer.kinematic_phase1.training.build_finisher_handoff_state_buffer import _finisher_ready  # noqa: E402


def main() -> None:
    policies, _ = gg.load_checkpoints()
    cfgs = gg.merged_configs()
    env_cfg = gg.to_env_config(cfgs["approach_dynamic_scale_big"])
    model = policies["approach_stage8_11"]
    seed, stage, n = 700001, 10, 48
    suite = gg.build_curriculum_local_eval_suite(env_cfg, seed=seed, stage_index=stage, n_episodes=n)
    out: dict[str, list] = {}
    put = lambda k, v: out.setdefault(k, []).append(v)  # noqa: E731
    for episode in suite:
        env = gg.ArmKinematicEnv(config=env_cfg)
        env.set_curriculum_stage(stage)
        env.set_policy_mode("approach")
        approach_result, first_handoff = gg._run_approach_with_handoff(
            env=env, model=model, reset_options={**episode.reset_options(), "policy_mode": "approach"}, ready_cfg=env_cfg.reward_config,
            handoff_confirm_steps=2)
        final_ready = _finisher_ready(approach_result, cfg=env_cfg.reward_config)
        opts = episode.reset_options()
        put("initial_q", np.asarray(opts["initial_q"], dtype=float))
        put("goal_q", np.asarray(opts["goal_q"], dtype=float))
        put("goal_pose6", np.asarray(opts["goal_pose6"], dtype=float))
        put("final_ready", int(final_ready))
        put("has_first", int(first_handoff is not None))
        for name, res in (("final", approach_result), ("first", first_handoff)):
            r = res if res is not None else {}
            put(f"{name}_step", int(r.get("step_count", -1)))
            put(f"{name}_q", np.asarray(r.get("final_q", np.zeros(7)), dtype=float))
            put(f"{name}_dq", np.asarray(r.get("final_dq", np.zeros(7)), dtype=float))
            put(f"{name}_prev_action", np.asarray(r.get("final_prev_action", np.zeros(7)), dtype=float))
            put(f"{name}_goal_q", np.asarray(r.get("goal_q", np.zeros(7)), dtype=float))
            put(f"{name}_goal_pose6", np.asarray(r.get("goal_pose6", np.zeros(6)), dtype=float))
            for k in ("final_position_error", "final_orientation_error", "final_action_magnitude", "final_dq_norm"):
                put(f"{name}_{k}", float(r.get(k, 0.0)))
    arrays = {k: np.asarray(v) for k, v in out.items()}
    np.savez_compressed(gg.GOLD / "handoff_states.npz", suite_seed=seed, stage_index=stage, dwell_steps_target=int(env_cfg.dwell_steps_target), **arrays)
    print("handoff_states.npz: final_ready", arrays["final_ready"].mean(), "first_confirmed", arrays["has_first"].mean())


if __name__ == "__main__":
    main()
