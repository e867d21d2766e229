"""From-scratch PPO on the kinematic Approach env with the curriculum shell: does the on-device trainer actually learn?

  python tools/learn_demo.py [--envs 16384] [--n-steps 128] [--iters 60] [--lr 3e-4]

A random-init MultiInputPolicy-shaped policy starts on Stage 0 of the official 12-stage table; the windowed success rate promotes the
stage (`PointCurriculumCallback` semantics).  Prints one JSON line per iteration (stage, rollout success rate, losses, env-steps so
far, wall seconds) and a final line with a Stage-0..k evaluation of the trained policy by the fused Approach-only rollout.
"""
import argparse, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from rl_brain_trainer_b200 import config as kcfg, gate, ppo

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=16384)
ap.add_argument("--n-steps", type=int, default=128)
ap.add_argument("--iters", type=int, default=60)
ap.add_argument("--lr", type=float, default=3e-4)
ap.add_argument("--epochs", type=int, default=8)
ap.add_argument("--minibatches", type=int, default=16)
ap.add_argument("--log-std", type=float, default=-0.5)
ap.add_argument("--ent", type=float, default=0.0)
ap.add_argument("--clip", type=float, default=0.2)
ap.add_argument("--gamma", type=float, default=0.98)
ap.add_argument("--from-checkpoint", action="store_true", help="fine-tune the bundled Approach checkpoint instead of a random init")
ap.add_argument("--stage", type=int, default=0, help="curriculum stage the training starts on")
a = ap.parse_args()

dev = torch.device("cuda", 0)
cfg = kcfg.load_preset("approach_dynamic_scale_big")
from rl_brain_trainer_b200.policy import PolicyWeights
pol = PolicyWeights.preset("approach_stage8_11", dev) if a.from_checkpoint else ppo.random_policy(56, seed=0, log_std_init=a.log_std, device=dev)
fcfg, fin = kcfg.load_preset("finisher_noop_ft"), PolicyWeights.preset("finisher", dev)


def evaluate(stages):
    ev = gate.evaluate_workspace_expansion(cfg, pol, fcfg, fin, episodes=2048, seed=720001, stage_indices=stages)
    return {str(s): round(ev["stage_metrics"][s]["success_rate"], 4) for s in stages}


if a.from_checkpoint:
    print(json.dumps({"before": True, "approach_finisher_success_by_stage": evaluate(sorted({0, 5, a.stage}))}), flush=True)
S = a.envs * a.n_steps
hp = ppo.PPOHyper(learning_rate=a.lr, n_steps=a.n_steps, batch_size=S // a.minibatches, n_epochs=a.epochs, gamma=a.gamma, gae_lambda=0.95, clip_range=a.clip,
                  ent_coef=a.ent)
tr = ppo.PPOTrainer(cfg, pol, num_envs=a.envs, hyper=hp, device=dev, seed=1, stage_index=a.stage)
t0 = time.perf_counter()
for it in range(a.iters):
    row = tr.learn(1)[0]
    print(json.dumps({"iter": it, "stage": int(row["stage"]), "episodes": int(row["episodes"]),
                      "rollout_success_rate": round(row["successes"] / max(row["episodes"], 1.0), 4), "mean_reward": round(row["mean_reward"], 4),
                      "value_loss": round(row["value_loss"], 4), "approx_kl": round(row["approx_kl"], 5), "log_std": round(float(pol.tensors["log_std"].mean()), 3), "env_steps": int(row["timesteps"]),
                      "wall_s": round(time.perf_counter() - t0, 2)}), flush=True)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
stages = sorted({0, 5, a.stage}) if a.from_checkpoint else list(range(0, min(int(tr.env.get_curriculum_stage()) + 2, 12)))
print(json.dumps({"final": True, "wall_s": round(wall, 2), "env_steps": int(tr.num_timesteps), "stage_reached": int(tr.env.get_curriculum_stage()),
                  "approach_finisher_success_by_stage": evaluate(stages)}))
