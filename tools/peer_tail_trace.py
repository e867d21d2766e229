"""Where the cycles of the fused exchange + Adam tail of kin_ppo_grad_tc_update go (debug build with -DKIN_PPO_TRACE, see ppo_trace.py --build).

  torchrun --nproc-per-node N tools/peer_tail_trace.py [--envs 65536]      (N = 1 works too: one-rank buffer)

Rank 0 prints, for thread 0 of CTAs 0 and 150, the mean cycles per launch of: grid barrier 1 (includes waiting for the slowest CTA's tiles),
slice reduction + stores to the peers, release flags, flag wait (the peers' delivery), rank-ordered sum, norm partial + grid barrier 2, Adam.
"""
import argparse, ctypes, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist

from rl_brain_trainer_b200 import _lib
_lib.LIB_PATH = ROOT / "tools" / "_trace_ppo" / "libkin_b200_trace.so"
from rl_brain_trainer_b200 import config as kcfg, ppo

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = kcfg.load_preset("approach_dynamic_scale_big")
pol = ppo.random_policy(56, seed=0, log_std_init=-1.0, device=dev)
S = a.envs * 128
hp = ppo.PPOHyper(learning_rate=4e-6, n_steps=128, batch_size=S // 16, n_epochs=8, gamma=0.995, gae_lambda=0.95, clip_range=0.1, ent_coef=0.0003)
tr = ppo.PPOTrainer(cfg, pol, num_envs=a.envs, hyper=hp, device=dev, seed=1, stage_index=10, process_group=dist.group.WORLD if world > 1 else None,
                    grad_exchange="peer" if world > 1 else "nccl", fused_update=True)
for _ in range(3):
    tr.collect(); tr.update()
torch.cuda.synchronize(dev)
if rank == 0:
    L = _lib.lib()
    buf = (ctypes.c_ulonglong * 24)()
    L.kin_debug_peer_trace.argtypes = [ctypes.c_void_p]
    assert L.kin_debug_peer_trace(buf) == 0
    names = ["grid barrier 1 (slowest CTA's tiles + barrier)", "slice reduction + stores to peers", "release flags (st.release.sys)",
             "flag wait (peers' delivery)", "rank-ordered sum -> grad", "sum of squares + grid barrier 2", "clip coefficient + Adam on the slice"]
    for c, who in enumerate(("CTA 0", "CTA 150")):
        row = [buf[c * 12 + i] for i in range(12)]
        n = max(row[11], 1)
        print(f"{who}: {n} launches, world {world}: {sum(row[:7]) / n:.0f} cycles per launch in the tail")
        for i, nm in enumerate(names):
            print(f"   {nm:50s} {row[i] / n:9.0f}")
if rank == 0 and hasattr(L, "kin_debug_peer_wait"):
    import numpy as np
    wbuf = (ctypes.c_ulonglong * 1024)()
    L.kin_debug_peer_wait.argtypes = [ctypes.c_void_p]
    assert L.kin_debug_peer_wait(wbuf) == 0
    w = np.array(list(wbuf), dtype=np.float64).reshape(512, 2)[:296]
    n = max(buf[11], 1)
    wait, sm = w[:, 0] / n, w[:, 1].astype(int)
    order = np.argsort(wait)
    print("barrier-1 wait per CTA (cycles per launch): min %.0f  median %.0f  max %.0f" % (wait.min(), np.median(wait), wait.max()))
    print("actor CTAs (0..147): mean wait %.0f; critic CTAs (148..295): mean wait %.0f" % (wait[:148].mean(), wait[148:].mean()))
    print("CTAs with 28 tiles (x < 100): mean wait %.0f; 27 tiles: %.0f" % (np.concatenate([wait[:100], wait[148:248]]).mean(), np.concatenate([wait[100:148], wait[248:]]).mean()))
    print("the 12 CTAs that wait least (= finish last): " + ", ".join(f"cta {c} sm {sm[c]} {wait[c]:.0f}" for c in order[:12]))
    per_sm = {}
    for c in range(296):
        per_sm.setdefault(sm[c], []).append(wait[c])
    sm_wait = sorted((min(v), k, len(v)) for k, v in per_sm.items())
    print("SMs whose CTAs finish last: " + ", ".join(f"sm {k} ({n_} CTAs) {w_:.0f}" for w_, k, n_ in sm_wait[:10]))
    print("SMs whose CTAs finish first: " + ", ".join(f"sm {k} ({n_} CTAs) {w_:.0f}" for w_, k, n_ in sm_wait[-6:]))
tr.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
