"""Time kin_step_kernel variants at an HBM-resident size (2 M envs): dock, approach, approach + auto-reset.

  [KIN_B200_LIB=<alternate build>] python tools/step_prof.py [--envs 2097152] [--launches 20]
"""
import argparse, json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from rl_brain_trainer_b200 import config as kcfg
from rl_brain_trainer_b200.env import BatchedArmKinematicEnv

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 21)
ap.add_argument("--launches", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(2)
act = torch.rand((a.envs, 7), device=dev, generator=g) * 2 - 1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(a.launches):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot * 1e3 / a.launches


out = {}
for name, preset, nbytes, kw in (("dock", "finisher_noop_ft", 548, {}), ("approach", "approach_dynamic_scale_big", 532, {}),
                                 ("approach_autoreset", "approach_dynamic_scale_big", 532, {"auto_reset": True})):
    env = BatchedArmKinematicEnv(kcfg.load_preset(preset), a.envs, dev, with_aux=False, seed=3, host_sampler=False, **kw)
    if kw:
        env.set_curriculum_stage(5)
    env.reset()
    us = timed(lambda: env.step_raw(act))
    out[name] = {"us_per_launch": round(us, 2), "gbs": round(nbytes * a.envs / us / 1e3, 1)}
    del env
print(json.dumps(out))
