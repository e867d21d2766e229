"""Which Stage-5 episodes flip between the GPU rollout variants and the fp64 oracle, and are they threshold-sensitive? (GPU)"""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from rl_brain_trainer_b200 import config as kcfg
from rl_brain_trainer_b200.policy import PolicyWeights
from rl_brain_trainer_b200.rollout import ApproachFinisherRollout
from rl_brain_trainer_b200.samplers import build_curriculum_local_eval_suite
from tests._util import oracle_policy, threshold_sensitive_episodes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
acfg, fcfg = kcfg.load_preset("approach_dynamic_scale_big"), kcfg.load_preset("finisher_noop_ft")
pa, pf = PolicyWeights.preset("approach_stage8_11"), PolicyWeights.preset("finisher")
suite = build_curriculum_local_eval_suite(acfg, seed=700001 + 5 * 1009, stage_index=5, n_episodes=n)
gpu = {name: ApproachFinisherRollout(acfg, pa, fcfg, pf, variant=v).evaluate_suite(suite).to_numpy() for name, v in (("strict", 0), ("tc", 1))}
A, F = oracle_policy("approach_stage8_11"), oracle_policy("finisher")
for band in (0.01, 0.02, 0.05, 0.10):
    t0 = time.time()
    ref, sens = threshold_sensitive_episodes(acfg, fcfg, A, F, suite, band, n_threads=os.cpu_count())
    line = f"band {band:.2f}: {int(sens.sum())} sensitive episodes of {n} ({time.time() - t0:.1f} s);"
    for name, r in gpu.items():
        flip = r["success"].astype(int) != ref["success"]
        line += f"  {name}: {int(flip.sum())} flips, {int((flip & ~sens).sum())} unexplained;"
    print(line)
for name, r in gpu.items():
    flip = np.nonzero(r["success"].astype(int) != ref["success"])[0]
    print(name, "flipped episodes:", flip[:40].tolist())
    for e in flip[:12]:
        print(f"   ep {e}: oracle succ {ref['success'][e]} kind {ref['handoff_kind'][e]} final pos {ref['final_position_error'][e]*1e3:.3f} mm ori {ref['final_orientation_error'][e]:.4f} "
              f"| gpu kind {r['handoff_kind'][e]} pos {r['final_position_error'][e]*1e3:.3f} ori {r['final_orientation_error'][e]:.4f} "
              f"| approach-final oracle pos {ref['approach_final_position_error'][e]*1e3:.3f} ori {ref['approach_final_orientation_error'][e]:.4f} streak {ref['max_ready_streak'][e]}")
