"""Diagnostic: SM-clock trace of the three GEMM round trips per env-step in kin_rollout_tc16_kernel (debug build, -DKIN_TC16_TRACE).

  python tools/tc16_trace.py --build     # here (nvcc): tools/_trace/libkin_b200_trace.so
  python tools/tc16_trace.py             # on a B200: per layer -- arrival skew within a tile, last arrival -> MMAs committed,
                                         # committed -> first wake, wake spread; and the time between the round trips
The trace build is a separate library; the product library never carries the stamps.
"""
import ctypes, os, subprocess, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "tools" / "_trace"
TRACE_LIB = OUT / "libkin_b200_trace.so"

if "--build" in sys.argv:
    from rl_brain_trainer_b200 import build as kb
    OUT.mkdir(exist_ok=True)
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    objs = []
    for src in sorted(kb.CSRC.glob("*.cu")):
        obj = OUT / f"trace_{src.stem}.o"
        subprocess.run([kb._nvcc(), *kb.NVCC_FLAGS, "-DKIN_TC16_TRACE", "-c", str(src), "-o", str(obj)], check=True, env=env)
        objs.append(str(obj))
    subprocess.run([kb._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(TRACE_LIB), *objs, "-lcudart"], check=True, env=env)
    print(TRACE_LIB)
    sys.exit(0)

import torch
from rl_brain_trainer_b200 import _lib
_lib.LIB_PATH = TRACE_LIB
from rl_brain_trainer_b200 import config as kcfg
from rl_brain_trainer_b200.policy import PolicyWeights
from rl_brain_trainer_b200.rollout import ApproachFinisherRollout
from rl_brain_trainer_b200.samplers import build_curriculum_local_eval_suite

n = 65536
acfg, fcfg = kcfg.load_preset("approach_dynamic_scale_big"), kcfg.load_preset("finisher_noop_ft")
pa, pf = PolicyWeights.preset("approach_stage8_11"), PolicyWeights.preset("finisher")
suite = build_curriculum_local_eval_suite(acfg, seed=700001 + 5 * 1009, stage_index=5, n_episodes=n)
ro = ApproachFinisherRollout(acfg, pa, fcfg, pf, variant=1)
dev = ro.upload(suite)
for _ in range(3):
    ro.run(dev)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (8 * 16 * 12))()
L = _lib.lib()
L.kin_debug_tc16_trace.argtypes = [ctypes.c_void_p]
assert L.kin_debug_tc16_trace(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(8, 16, 12)
n_warps = int((t[0, :, 0] != 0).sum())
print(f"{n_warps} warps in CTA 0")
rows = {k: [] for k in ("skew", "issue", "iss_wake", "mma", "wake_spread", "wait_mean", "phase")}
for s in range(8):
    for tile in range((n_warps + 3) // 4):
        w = slice(4 * tile, min(4 * tile + 4, n_warps))
        for layer in range(3):
            arr, wake, com = t[s, w, layer], t[s, w, 3 + layer], t[s, w, 6 + layer]
            c = com.max()   # only the issuing warp stamps this slot in this step; stale stamps of other warps are older
            rows["skew"].append(arr.max() - arr.min())
            rows["issue"].append(c - arr.max())
            rows["iss_wake"].append(t[s, w, 9 + layer].max() - arr.max())
            rows["mma"].append(wake.min() - c)
            rows["wake_spread"].append(wake.max() - wake.min())
            rows["wait_mean"].append((wake - arr).mean())
        if s + 1 < 8:
            rows["phase"].append([(t[s, w, 1] - t[s, w, 3]).mean(), (t[s, w, 2] - t[s, w, 4]).mean(), (t[s + 1, w, 0] - t[s, w, 5]).mean()])
for k in ("skew", "issue", "iss_wake", "mma", "wake_spread", "wait_mean"):
    a = np.array(rows[k]).reshape(-1, 3)
    print(f"{k:12s} L1 {a[:, 0].mean():7.0f}  L2 {a[:, 1].mean():7.0f}  L3 {a[:, 2].mean():7.0f}   (cycles, mean over tiles x steps; max {a.max():.0f})")
ph = np.array(rows["phase"])
print(f"compute      epi1 {ph[:, 0].mean():7.0f}  epi2 {ph[:, 1].mean():7.0f}  act+env+obs {ph[:, 2].mean():7.0f}")
step = (t[1:, :n_warps, 0] - t[:-1, :n_warps, 0]).mean()
print(f"cycles per env-step (warp mean): {step:.0f}")
