"""Diagnostic: where the cycles of one 128-sample tile of kin_ppo_grad_tc go (debug build with -DKIN_PPO_TRACE).

  python tools/ppo_trace.py --build        # here (nvcc, no GPU needed): tools/_trace_ppo/libkin_b200_trace.so
  python tools/ppo_trace.py [--envs 65536] # on a B200: one PPO update, then the per-phase cycle table

The trace build is a separate library; the product library never carries the counters.
"""
import argparse, ctypes, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "tools" / "_trace_ppo"
TRACE_LIB = OUT / "libkin_b200_trace.so"

ap = argparse.ArgumentParser()
ap.add_argument("--build", action="store_true")
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--n-steps", type=int, default=128)
a = ap.parse_args()

if a.build:
    from rl_brain_trainer_b200 import build as kb
    OUT.mkdir(exist_ok=True)
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    objs = []
    for src in sorted(kb.CSRC.glob("*.cu")):
        obj = OUT / f"trace_{src.stem}.o"
        subprocess.run([kb._nvcc(), *kb.NVCC_FLAGS, "-DKIN_PPO_TRACE", "-c", str(src), "-o", str(obj)], check=True, env=env)
        objs.append(str(obj))
    subprocess.run([kb._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(TRACE_LIB), *objs, "-lcudart"], check=True, env=env)
    print(TRACE_LIB)
    sys.exit(0)

import torch
from rl_brain_trainer_b200 import _lib
_lib.LIB_PATH = TRACE_LIB
from rl_brain_trainer_b200 import config as kcfg, ppo

dev = torch.device("cuda", 0)
cfg = kcfg.load_preset("approach_dynamic_scale_big")
pol = ppo.random_policy(56, seed=0, log_std_init=-1.0, device=dev)
S = a.envs * a.n_steps
hp = ppo.PPOHyper(learning_rate=4e-6, n_steps=a.n_steps, batch_size=S // 16, n_epochs=1, gamma=0.995, gae_lambda=0.95, clip_range=0.1, ent_coef=0.0003)
tr = ppo.PPOTrainer(cfg, pol, num_envs=a.envs, hyper=hp, device=dev, seed=1, stage_index=10, update_variant="tc")
tr.collect(); tr.update(); tr.collect(); tr.update()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 64)()
L = _lib.lib()
L.kin_debug_ppo_trace.argtypes = [ctypes.c_void_p]
assert L.kin_debug_ppo_trace(buf) == 0
epi_names = ["loss-input loads (+ X conversion)", "L1 wait", "epi1 + arrive", "L2 wait", "epi2 + arrive", "L3 wait", "loss + arrive", "bwd1 wait",
             "G2 math + ride wait", "G2 store + arrive, bwd2 wait", "G1 math + ride wait", "G1 store + arrive", "loop top"]
iss_names = ["X wait", "L1 issue", "wg wait + prefetch", "wait H1", "wait H2 (+ L2 issue)", "wait dO (+ L3 issue)", "wait G2 (+ bwd1 issue)",
             "wait G1 (+ bwd2 issue)", "trailing issue", "-", "-", "-", "loop top"]
if hasattr(L, "kin_debug_ppo_trace3") and os.environ.get("KIN_PPO_TC3", "1") != "0":      # the three-stream kernel ran instead
    L.kin_debug_ppo_trace3.argtypes = [ctypes.c_void_p]
    assert L.kin_debug_ppo_trace3(buf) == 0
    epi3 = ["L1 wait", "epi1 math", "wg wait + H1 store + arrive", "L2 wait", "epi2 + arrive", "L3 wait", "loss + arrive", "bwd1 wait", "G2 math",
            "ride wait", "G2 store + arrive", "bwd2 wait", "G1 math", "ride wait + G1 store + arrive", "loop top"]
    chain3 = ["X wait", "L1 issue", "H1 wait", "L2 issue + H2 wait", "L3 issue + dO wait", "bwd1 issue + G2 wait", "bwd2 issue + G1 wait", "loop top"] + ["-"] * 7
    acc3 = ["-", "-", "-", "dWO batch + prefetches", "dW1 batch", "dW0 batch", "polling (nothing ready / not its turn)"] + ["-"] * 8
    names = [chain3, epi3, acc3, epi3]
    for w, name in enumerate(["first CTA (actor) chain issuer of stream 0", "first CTA (actor) stream 0 tid32", "first CTA (actor) accumulate warp, all 3 streams",
                              "last CTA (critic) stream 0 tid32"]):
        row = [buf[w * 16 + i] for i in range(16)]
        n = max(row[15], 1)
        print(f"{name}: {n} tiles (stream 0), {sum(row[:15]) / n:.0f} cycles per stream-0 tile")
        for i in range(15):
            print(f"   {names[w][i]:40s} {row[i] / n:8.0f}")
    sys.exit(0)
who = ["cta(0,0) issuer", "cta(0,0) tid32", "cta(1,1) issuer", "cta(1,1) tid32"]
for w in range(4):
    row = [buf[w * 16 + i] for i in range(16)]
    n = max(row[15], 1)
    print(f"{who[w]}: {n} tiles, {sum(row[:13]) / n:.0f} cycles/tile")
    for i in range(13):
        print(f"   {(epi_names if w & 1 else iss_names)[i]:40s} {row[i] / n:8.0f}")
