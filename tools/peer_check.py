"""Multi-GPU check of the NVLink peer-memory gradient exchange against the NCCL all-reduce.

  torchrun --nproc-per-node 2 tools/peer_check.py

Two trainers per rank from the same seeds, one per exchange; after two collect + update iterations the parameters must agree to
float rounding (the two exchanges add the ranks in different orders) and the peer path's parameters must be bitwise identical
on every rank.  Prints one JSON line on rank 0.
"""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist

from rl_brain_trainer_b200 import config as kcfg, ppo

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=dev)
cfg = kcfg.load_preset("approach_dynamic_scale_big")
N, T = 4096, 32
hp = ppo.PPOHyper(learning_rate=3e-4, n_steps=T, batch_size=N * T // 4, n_epochs=2, gamma=0.995, gae_lambda=0.95, clip_range=0.1, ent_coef=0.0003)
out = {}
params = {}
# "peer": the default -- one launch per minibatch (exchange + clip + Adam in the gradient kernel's tail); "peer_separate_adam": the same
# in-kernel exchange followed by a kin_ppo_adam launch; "peer_two_kernel": push + gather kernels, then kin_ppo_adam
for ex in ("nccl", "peer", "peer_separate_adam", "peer_two_kernel"):
    pol = ppo.random_policy(56, seed=0, log_std_init=-1.0, device=dev)
    tr = ppo.PPOTrainer(cfg, pol, num_envs=N, hyper=hp, device=dev, seed=1, stage_index=10, update_variant="tc", grad_exchange=ex.split("_")[0],
                        fused_update=None if ex in ("nccl", "peer") else False)
    tr.fused_exchange = not ex.endswith("two_kernel")
    for _ in range(2):
        tr.collect()
        u = tr.update()
    torch.cuda.synchronize(dev)
    params[ex] = tr.params.clone()
    out[ex] = {"approx_kl": u["approx_kl"], "grad_norm": u["grad_norm"], "value_loss": u["value_loss"]}
    tr.close()
rel = float((params["peer"] - params["nccl"]).norm() / params["nccl"].norm())
# the in-kernel exchange sums in the two-kernel form's orders: bitwise equal with the same (separate) Adam; the one-launch form adds the
# squares for the clip norm in another order (a few ulp of the coefficient)
fused_equal = bool(torch.equal(params["peer_separate_adam"], params["peer_two_kernel"]))
rel_one_launch = float((params["peer"] - params["peer_two_kernel"]).norm() / params["peer_two_kernel"].norm())
gathered = [torch.zeros_like(params["peer"]) for _ in range(world)]
dist.all_gather(gathered, params["peer"])
identical = all(bool(torch.equal(gathered[0], g)) for g in gathered)
if rank == 0:
    print(json.dumps({"world": world, "peer_vs_nccl_rel_diff": rel, "peer_params_bitwise_identical_across_ranks": identical,
                      "fused_tail_bitwise_equals_two_kernel": fused_equal, "one_launch_vs_two_kernel_rel_diff": rel_one_launch, **out}))
dist.destroy_process_group()
assert rel < 1e-5 and identical and fused_equal and rel_one_launch < 1e-6
