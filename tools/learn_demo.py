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
ap.add_argument("--target-kl", type=float, default=None)
ap.add_argument("--window", type=int, default=None, help="curriculum window_episodes override (the reference's 32 assumes 16 envs)")
ap.add_argument("--min-episodes", type=int, default=None)
ap.add_argument("--lr-final", type=float, default=None, help="linear learning-rate decay to this value over the run")
ap.add_argument("--oracle-eval", action="store_true", help="evaluate the final policy with the fp64 CPU oracle too (tests-only code, used here as the judge)")
ap.add_argument("--save", default=None)
ap.add_argument("--preset", default="approach_dynamic_scale_big", help="env preset; approach_default is the reference's from-scratch config (train_approach_policy.py)")
ap.add_argument("--approach-only", action="store_true", help="evaluate the Approach policy alone (its env's own success criterion), no Finisher handoff")
ap.add_argument("--stop-at-success", type=float, default=None, help="stop once the LAST curriculum stage is reached and a rollout's success rate is at least this "
                "(best-checkpoint selection in the spirit of the reference's eval-gate callback: longer training trades success for shaped reward)")
ap.add_argument("--gate-every", type=int, default=0, help="evaluate with the reference's gate (gated_score / retention_ok) every this many iterations and "
                "finish on the best checkpoint (WorkspaceEvalGateCallback, train_workspace_expansion.py:54-129)")
ap.add_argument("--load", default=None, help="start from a checkpoint written by --save")
ap.add_argument("--shuffle", default="tile", choices=("tile", "sample", "sample_once"), help="minibatch composition (sample = SB3's per-sample permutation every epoch)")
a = ap.parse_args()

dev = torch.device("cuda", 0)
cfg = kcfg.load_preset(a.preset)
if a.window or a.min_episodes:
    from dataclasses import replace
    cur = cfg.curriculum_config
    cfg = replace(cfg, curriculum_config=replace(cur, window_episodes=a.window or cur.window_episodes, min_episodes_per_stage=a.min_episodes or cur.min_episodes_per_stage))
from rl_brain_trainer_b200.policy import PolicyWeights
pol = PolicyWeights.preset("approach_stage8_11", dev) if a.from_checkpoint else ppo.random_policy(56, seed=0, log_std_init=a.log_std, device=dev)
fcfg, fin = kcfg.load_preset("finisher_noop_ft"), PolicyWeights.preset("finisher", dev)


if a.load:
    pol = PolicyWeights.load(a.load, dev)


def evaluate(stages):
    ev = gate.evaluate_workspace_expansion(cfg, pol, None if a.approach_only else fcfg, None if a.approach_only else fin, episodes=2048, seed=720001, stage_indices=stages)
    return {str(s): round(ev["stage_metrics"][s]["success_rate"], 4) for s in stages}


def oracle_eval(stages):
    """The fp64 CPU oracle as the judge of a trained policy (test infrastructure; this tool is not part of the product path)."""
    import numpy as np
    from oracle import kin_oracle as ko
    from rl_brain_trainer_b200.samplers import build_curriculum_local_eval_suite
    sd = {k: v.detach().cpu().numpy() for k, v in pol.state_dict().items()}
    fw = dict(np.load(kcfg.PRESET_DIR / "policies" / "finisher.npz"))
    oe = {}
    for st in stages:
        suite = build_curriculum_local_eval_suite(cfg, seed=700001 + st * 1009, stage_index=st, n_episodes=2048)
        ref, _ = ko.eval_approach_finisher(ko.params_from_config(cfg), None if a.approach_only else ko.params_from_config(fcfg), ko.OracleMlp(sd),
                                           None if a.approach_only else ko.OracleMlp(fw),
                                           initial_q=suite.initial_q.astype(np.float32).astype(float), goal_q=suite.goal_q.astype(np.float32).astype(float), n_threads=16)
        oe[str(st)] = round(float(ref["success"].mean()), 4)
    return oe


if a.from_checkpoint:
    if a.oracle_eval:
        print(json.dumps({"before": True, "oracle_approach_finisher_success_by_stage": oracle_eval(sorted({0, 5, a.stage}))}), flush=True)
    print(json.dumps({"before": True, "approach_finisher_success_by_stage": evaluate(sorted({0, 5, a.stage}))}), flush=True)
S = a.envs * a.n_steps
hp = ppo.PPOHyper(learning_rate=a.lr, n_steps=a.n_steps, batch_size=S // a.minibatches, n_epochs=a.epochs, gamma=a.gamma, gae_lambda=0.95, clip_range=a.clip,
                  ent_coef=a.ent, target_kl=a.target_kl)
tr = ppo.PPOTrainer(cfg, pol, num_envs=a.envs, hyper=hp, device=dev, seed=1, stage_index=a.stage, shuffle=a.shuffle)
eg = None
if a.gate_every:
    we = kcfg.preset_dict(a.preset).get("workspace_expansion", {})
    eg = gate.EvalGate(cfg, fcfg, fin, eval_interval=a.gate_every * S, episodes=2048, seed=int(we.get("gate_suite_seed", 720001)),
                       stage_indices=sorted({0, 5, 8, a.stage}), gate_config={**we.get("gate", {}), "score_stage_index": a.stage})
    print(json.dumps({"gate_initial": eg.maybe_eval(0, pol, force=True)}), flush=True)      # the starting policy is a candidate too
t0 = time.perf_counter()
for it in range(a.iters):
    if a.lr_final is not None:
        tr.hp.learning_rate = a.lr + (a.lr_final - a.lr) * it / max(a.iters - 1, 1)
    row = tr.learn(1, gate=eg)[0]
    print(json.dumps({"iter": it, "stage": int(row["stage"]), "episodes": int(row["episodes"]),
                      "rollout_success_rate": round(row["successes"] / max(row["episodes"], 1.0), 4), "mean_reward": round(row["mean_reward"], 4),
                      "minibatches": int(row["minibatches"]), "value_loss": round(row["value_loss"], 4), "approx_kl": round(row["approx_kl"], 5), "log_std": round(float(pol.tensors["log_std"].mean()), 3), "env_steps": int(row["timesteps"]),
                      "wall_s": round(time.perf_counter() - t0, 2), **({"gate_score": round(row["gate_score"], 4)} if "gate_score" in row else {})}), flush=True)
    if a.stop_at_success is not None and int(row["stage"]) >= len(cfg.curriculum_config.stages) - 1 and row["successes"] / max(row["episodes"], 1.0) >= a.stop_at_success:
        break
torch.cuda.synchronize()
wall = time.perf_counter() - t0
if eg is not None and eg.best_state is not None:      # finish on the gate's best checkpoint
    for f, key in ppo.KEYS.items():
        if f in pol.tensors and key in eg.best_state:
            pol.tensors[f].copy_(eg.best_state[key])
    print(json.dumps({"gate_history": eg.history, "best_score": eg.best_score}), flush=True)
n_stages = len(cfg.curriculum_config.stages)
stages = sorted({0, 5, a.stage}) if a.from_checkpoint else list(range(0, min(int(tr.env.get_curriculum_stage()) + 2, n_stages)))
extra = {}
if a.oracle_eval:
    extra["oracle_approach_finisher_success_by_stage"] = oracle_eval(stages)
if a.save:
    tr.save_checkpoint(a.save)
print(json.dumps({"final": True, **extra, "wall_s": round(wall, 2), "env_steps": int(tr.num_timesteps), "stage_reached": int(tr.env.get_curriculum_stage()),
                  "approach_finisher_success_by_stage": evaluate(stages)}))
