"""Config 5 of BASELINE.json: Stage 10/11 stress-shell PPO training, rollout + actor/critic update (+ NCCL grad all-reduce).

  python tools/train_bench.py [--envs 16384] [--n-steps 128] [--epochs 8] [--iters 3]
  torchrun --nproc-per-node N tools/train_bench.py ...        (one rank per GPU)

Prints one JSON line: rollout / update / end-to-end env-steps/s (whole job), timed with CUDA events, max over ranks.
"""
import argparse, json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist

from rl_brain_trainer_b200 import config as kcfg, ppo
from rl_brain_trainer_b200.policy import PolicyWeights

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=16384)
ap.add_argument("--n-steps", type=int, default=128)
ap.add_argument("--epochs", type=int, default=8)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--stage", type=int, default=10)
ap.add_argument("--update", default="tc", choices=("tc", "fp32"))
ap.add_argument("--grad-exchange", default="peer", choices=("peer", "nccl"), help="per-minibatch gradient sum over ranks: NVLink peer buffers or NCCL")
ap.add_argument("--force-peer", action="store_true", help="use the peer-memory exchange even with one rank (it then pushes into its own buffer)")
ap.add_argument("--two-kernel", action="store_true", help="peer exchange as separate push + gather kernels instead of the gradient kernel's fused tail")
ap.add_argument("--route", action="store_true", help="train the 80-input route policy on the batched RouteSequence env (train_route_curriculum.py)")
ap.add_argument("--fused-update", default="auto", choices=("auto", "on", "off"), help="kin_ppo_grad_tc_update: gradient + reduction + exchange + clip + Adam in one launch per minibatch (auto: with several ranks on the peer exchange)")
ap.add_argument("--shuffle", default="tile", choices=("tile", "sample", "sample_once"), help="minibatch composition: tile unions, SB3's per-sample permutation every epoch, or one per rollout")
ap.add_argument("--from-checkpoint", action="store_true", help="fine-tune the bundled approach checkpoint instead of a random init")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
route_kw = {}
if a.route:
    from rl_brain_trainer_b200.route import synthetic_route
    route = synthetic_route(483, seed=7)
    cfg, seq = kcfg.to_route_env_config(kcfg.preset_dict("route_prefix120"), max_route_index=len(route) - 1)
    route_kw = dict(route=route, route_sequence_config=seq)
    pol = PolicyWeights.preset("route_prefix120", dev) if a.from_checkpoint else ppo.random_policy(80, seed=0, log_std_init=-1.0, device=dev)
else:
    cfg = kcfg.load_preset("approach_dynamic_scale_big")
    pol = PolicyWeights.preset("approach_stage8_11", dev) if a.from_checkpoint else ppo.random_policy(56, seed=0, log_std_init=-1.0, device=dev)
S = a.envs * a.n_steps
hp = ppo.PPOHyper(learning_rate=4e-6, n_steps=a.n_steps, batch_size=S // 16, n_epochs=a.epochs, gamma=0.995, gae_lambda=0.95, clip_range=0.1, ent_coef=0.0003)
tr = ppo.PPOTrainer(cfg, pol, num_envs=a.envs, hyper=hp, device=dev, seed=1, stage_index=a.stage, update_variant=a.update,
                    grad_exchange=a.grad_exchange if (world > 1 or a.force_peer) else "nccl", shuffle=a.shuffle,
                    fused_update={"auto": None, "on": True, "off": False}[a.fused_update], **route_kw)
if a.two_kernel:
    tr.fused_exchange = False
tr.collect(); tr.update()          # warm-up
torch.cuda.synchronize(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
t_roll = t_upd = 0.0
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(a.iters):
    e0, e1, e2 = ev(), ev(), ev()
    e0.record(); r = tr.collect(); e1.record(); u = tr.update(); e2.record()
    torch.cuda.synchronize(dev)
    t_roll += e0.elapsed_time(e1) * 1e-3
    t_upd += e1.elapsed_time(e2) * 1e-3
wall = time.perf_counter() - t0
t = torch.tensor([t_roll, t_upd, wall], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
steps = S * a.iters * world
if rank == 0:
    print(json.dumps({"workload": "route_prefix120_ppo_train" if a.route else f"stage{a.stage}_ppo_train", "n_gpus": world, "envs_per_gpu": a.envs, "n_steps": a.n_steps, "epochs": a.epochs, "update_variant": a.update, "grad_exchange": a.grad_exchange if world > 1 else "none",
                      "minibatches_per_epoch": 16, "iters": a.iters, "rollout_env_steps_per_s": steps / float(t[0]),
                      "update_env_steps_per_s": steps / float(t[1]), "e2e_env_steps_per_s": steps / float(t[0] + t[1]),
                      "wall_env_steps_per_s": steps / float(t[2]), "rollout_s": float(t[0]) / a.iters, "update_s": float(t[1]) / a.iters,
                      "last": {**r, **u}}))
tr.close()
if world > 1:
    dist.destroy_process_group()
