"""One fused-rollout variant on the Stage-5 suite, CUDA-event timed (GPU).  usage: rollout_prof.py [n] [variant] [launches]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from rl_brain_trainer_b200 import config as kcfg
from rl_brain_trainer_b200.policy import PolicyWeights
from rl_brain_trainer_b200.rollout import ApproachFinisherRollout
from rl_brain_trainer_b200.samplers import build_curriculum_local_eval_suite

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 1
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 10
acfg, fcfg = kcfg.load_preset("approach_dynamic_scale_big"), kcfg.load_preset("finisher_noop_ft")
pa, pf = PolicyWeights.preset("approach_stage8_11"), PolicyWeights.preset("finisher")
suite = build_curriculum_local_eval_suite(acfg, seed=700001 + 5 * 1009, stage_index=5, n_episodes=n)
ro = ApproachFinisherRollout(acfg, pa, fcfg, pf, variant=variant)
dev = ro.upload(suite)
r = ro.run(dev)
torch.cuda.synchronize()
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(launches)]
for a, b in ev:
    flush.fill_(0.0)
    a.record()
    r = ro.run(dev)
    b.record()
torch.cuda.synchronize()
ms = np.array([a.elapsed_time(b) for a, b in ev])
res = r.to_numpy()
steps = int(res["approach_steps"].sum() + res["finisher_steps"].sum())
print(f"variant {variant} n {n}: {ms.mean():.4f} ms (min {ms.min():.4f}), {steps / ms.mean() / 1e6:.3f} G env-steps/s, success {res['success'].mean():.4f}")
