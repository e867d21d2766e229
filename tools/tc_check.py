"""Compare the tensor-core rollout variant against the strict-fp32 FFMA variant on the Stage-5 suite (GPU)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from rl_brain_trainer_b200 import config as kcfg
from rl_brain_trainer_b200.policy import PolicyWeights
from rl_brain_trainer_b200.rollout import ApproachFinisherRollout
from rl_brain_trainer_b200.samplers import build_curriculum_local_eval_suite

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
acfg, fcfg = kcfg.load_preset("approach_dynamic_scale_big"), kcfg.load_preset("finisher_noop_ft")
pa, pf = PolicyWeights.preset("approach_stage8_11"), PolicyWeights.preset("finisher")
suite = build_curriculum_local_eval_suite(acfg, seed=700001 + 5 * 1009, stage_index=5, n_episodes=n)
out = {}
for name, variant in (("ffma", 0), ("tc", 1)):
    ro = ApproachFinisherRollout(acfg, pa, fcfg, pf, variant=variant)
    dev = ro.upload(suite)
    r = ro.run(dev); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        r = ro.run(dev)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    out[name] = r.to_numpy()
    steps = int(out[name]["approach_steps"].sum() + out[name]["finisher_steps"].sum())
    print(f"{name}: {dt*1e3:.3f} ms/pass, {steps/dt/1e9:.3f} G env-steps/s, success {out[name]['success'].mean():.4f}, "
          f"final pos {out[name]['final_position_error'].mean()*1e3:.4f} mm ori {out[name]['final_orientation_error'].mean():.5f}")
for other in ("tc",):
  a, b = out["ffma"], out[other]
  print("----", other, "vs ffma")
  print("success flips:", int((a["success"] != b["success"]).sum()), "of", n)
  print("handoff kind diff:", int((a["handoff_kind"] != b["handoff_kind"]).sum()))
  print("max |final_q diff|:", float(np.abs(a["final_q"] - b["final_q"]).max()), "mean:", float(np.abs(a["final_q"] - b["final_q"]).mean()))
  print("approach pos err mean diff:", float(np.abs(a["approach_final_position_error"] - b["approach_final_position_error"]).mean()))
