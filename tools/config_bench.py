"""BASELINE.json configs 3 and 4 at full size (config 2 is bench.py, config 5 is tools/train_bench.py).

  python tools/config_bench.py randomstart [--pairs 1000000]      # mixed random-start known-workspace eval, sharded over the ranks
  python tools/config_bench.py route [--replicas 262144] [--end 170]   # dense holder-route sequential probe
  torchrun --nproc-per-node N tools/config_bench.py randomstart ...     # one rank per GPU, NCCL reduction of the statistics

Prints one JSON line per run: env-steps/s (CUDA events, max over ranks) and the success statistics.  The agreement of these configs
with the CPU oracle is checked in tests/ (test_gpu_rollout.py, test_gpu_route.py, test_gpu_fullsize.py), not here.
"""
import argparse, json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist

from rl_brain_trainer_b200 import config as kcfg, workspace as ws
from rl_brain_trainer_b200.distributed import allreduce_max_, allreduce_sum_, shard_slice
from rl_brain_trainer_b200.policy import PolicyWeights
from rl_brain_trainer_b200.rollout import ApproachFinisherRollout
from rl_brain_trainer_b200.samplers import EvalSuite

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["randomstart", "route"])
ap.add_argument("--pairs", type=int, default=1_000_000)
ap.add_argument("--replicas", type=int, default=262_144)
ap.add_argument("--end", type=int, default=170)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--oracle-sample", type=int, default=0, help="ignored (the oracle checks of these configs live in tests/)")
ap.add_argument("--variant", default="tc", choices=["tc", "ffma"], help="randomstart: tensor-core (tf32 MLP) or strict-fp32 rollout kernel")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
group = dist.group.WORLD if world > 1 else None
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731


def timed(fn, iters):
    fn()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3 / iters], dtype=torch.float64, device=dev)
    allreduce_max_(t, group)
    return out, float(t)


if a.what == "randomstart":
    # eval_full_workspace_coverage.py:193-256 with the overnight random-start config: maps of seed 940001, pairs drawn with
    # replacement, the KNOWN split filter (target stage <= 8, class in {retention, local, medium}), 160 + 36 steps per episode
    acfg, fcfg = kcfg.load_preset("randomstart_overnight"), kcfg.load_preset("finisher_noop_ft")
    seed = 940001
    targets = ws.generate_workspace_target_map(acfg, seed=seed + 1, stage_samples_per_stage=96, random_samples=384)
    starts = ws.generate_workspace_start_state_map(acfg, seed=seed + 2, stage_samples_per_stage=48, random_samples=384)
    t0 = time.perf_counter()
    pairs = ws.build_pair_table(starts, targets, seed=seed + 3, pair_count=int(a.pairs * 3.1))     # ~34 % of random pairs pass the KNOWN filter
    tstage0 = np.where(targets.stage[pairs.target] < 0, 0, targets.stage[pairs.target])
    pool = np.nonzero((tstage0 <= 8) & np.isin(pairs.klass, (0, 1, 2)))[0][: a.pairs]
    host_s = time.perf_counter() - t0
    n_total = int(pool.size)
    mine = pool[shard_slice(n_total, rank, world)]
    suite = ws.pairs_to_suite(starts, targets, pairs, mine)
    pa, pf = PolicyWeights.preset("randomstart", dev), PolicyWeights.preset("finisher", dev)
    from rl_brain_trainer_b200.rollout import VARIANT_FFMA, VARIANT_TC
    ro = ApproachFinisherRollout(acfg, pa, fcfg, pf, device=dev, variant=VARIANT_TC if a.variant == "tc" else VARIANT_FFMA)
    d = ro.upload(suite)
    res, secs = timed(lambda: ro.run(d), a.iters)
    stats = torch.stack([res.success.double().sum(), torch.tensor(float(mine.size), dtype=torch.float64, device=dev),
                         res.final_position_error.double().sum(), res.final_orientation_error.double().sum(), res.env_steps.double().sum()])
    allreduce_sum_(stats, group)
    klass = torch.as_tensor(pairs.klass[mine], device=dev)
    per_class = torch.stack([torch.stack([(res.success.bool() & (klass == k)).sum(), (klass == k).sum()]) for k in range(3)]).double()
    allreduce_sum_(per_class, group)
    line = {"workload": "mixed_randomstart_known_split", "n_gpus": world, "pairs_total": n_total, "pairs_per_gpu": int(mine.size),
            "env_steps_per_s": float(stats[4]) / secs, "s_per_pass": secs, "env_steps": float(stats[4]),
            "success_rate": float(stats[0] / stats[1]), "mean_final_pos_err_m": float(stats[2] / stats[1]),
            "mean_final_ori_err_rad": float(stats[3] / stats[1]),
            "success_by_class": {name: float(per_class[k, 0] / max(per_class[k, 1], 1)) for k, name in enumerate(("retention", "local", "medium"))},
            "host_pair_table_s": host_s, "rollout_variant": a.variant, "reference_published_known_success_96_episodes": 0.802}
else:
    from rl_brain_trainer_b200.route import evaluate_sequential_route, synthetic_route

    route = synthetic_route(483, seed=7)            # the reference's 483-waypoint file is absent upstream
    renv, _seq = kcfg.to_route_env_config(kcfg.preset_dict("route_prefix120"), max_route_index=len(route) - 1)
    pol = PolicyWeights.preset("route_prefix120", dev)
    n = a.replicas // world
    out, secs = timed(lambda: evaluate_sequential_route(route, renv, pol, n_replicas=n, start_index=1, end_index=a.end, start_q_noise_std=0.0008,
                                                        seed=11 + rank, device=dev, variant="tc" if a.variant == "tc" else "fp32"), a.iters)
    hist = out["prefix_histogram"].double()
    steps = out["env_steps"].double()
    allreduce_sum_(hist, group)
    allreduce_sum_(steps, group)
    prefix = out["longest_success_prefix"].double()
    line = {"workload": f"dense_route_sequential_probe_to_{a.end}", "n_gpus": world, "replicas_total": n * world, "replicas_per_gpu": n,
            "waypoints_probed": a.end, "env_steps_per_s": float(steps) / secs, "s_per_pass": secs, "env_steps": float(steps),
            "mean_longest_success_prefix": float((hist * torch.arange(hist.numel(), device=dev)).sum() / hist.sum()),
            "full_prefix_fraction": float(hist[-1] / hist.sum()), "replica0_prefix": int(out["replica0_longest_success_prefix"]),
            "rank0_prefix_min_max": [float(prefix.min()), float(prefix.max())], "route": "synthetic 483 waypoints (seed 7)",
            "probe_variant": "tc" if a.variant == "tc" else "fp32"}
if rank == 0:
    print(json.dumps(line))
if world > 1:
    dist.destroy_process_group()
