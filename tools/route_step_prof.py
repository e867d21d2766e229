"""Launch kin_route_step_kernel at an HBM-resident size (for ncu / timing): 1 M replicas of the synthetic 483-waypoint route.

  python tools/route_step_prof.py [--envs 1048576] [--scale 0.1] [--steps 12] [--sequence]
"""
import argparse, json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from rl_brain_trainer_b200 import config as kcfg
from rl_brain_trainer_b200.route import BatchedRouteKinematicEnv, synthetic_route

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--scale", type=float, default=0.1)
ap.add_argument("--steps", type=int, default=12)
a = ap.parse_args()
dev = torch.device("cuda", 0)
route = synthetic_route(483, seed=7)
renv, _ = kcfg.to_route_env_config(kcfg.preset_dict("route_prefix120"), max_route_index=len(route) - 1)
env = BatchedRouteKinematicEnv(route, renv, a.envs, dev)
env.reset(seed=5)
g = torch.Generator(device=dev); g.manual_seed(2)
act = ((torch.rand((a.envs, 7), device=dev, generator=g) * 2 - 1) * a.scale).contiguous()
for _ in range(3):
    env.step_raw(act)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    env.step_raw(act)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / a.steps
print(json.dumps({"kernel": "kin_route_step_kernel", "envs": a.envs, "action_scale": a.scale, "us_per_launch": us, "env_steps_per_s": a.envs / us * 1e6,
                  "gbs_algorithmic_644B": 644 * a.envs / us / 1e3}))
