"""Diagnostic: tensor-core PPO gradient (kin_ppo_grad_tc) vs the strict-fp32 kernel and fp32 autograd, per tensor; timing of both.

  python tools/ppo_tc_check.py
"""
import ctypes, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from rl_brain_trainer_b200 import _lib, ppo
from tests.test_gpu_ppo import _setup, _torch_forward, _torch_ppo_loss

ppo_mod, pol, flat = _setup(seed=4)
import os
hp = ppo.PPOHyper(clip_range=float(os.environ.get("CLIP", "0.15")), ent_coef=0.01, vf_coef=0.5, normalize_advantage=True)
S = 64 * 4096
g = torch.Generator(device="cuda").manual_seed(2)
obs = (torch.rand((S, 56), device="cuda", generator=g) * 2 - 1).contiguous()
with torch.no_grad():
    mean, value = _torch_forward(pol, obs)
sigma = pol.tensors["log_std"].exp()
act = (mean + sigma * torch.randn((S, 7), device="cuda", generator=g)).contiguous()
exact_logp = torch.distributions.Normal(mean, sigma).log_prob(act).sum(-1)
old_logp = (exact_logp + 0.3 * torch.randn(S, device="cuda", generator=g)).contiguous()
adv = torch.randn(S, device="cuda", generator=g).contiguous()
ret = (value + torch.randn(S, device="cuda", generator=g)).contiguous()
sums = torch.stack([adv.reshape(-1, 64).double().sum(1), (adv.reshape(-1, 64).double() ** 2).sum(1)], dim=1).contiguous()
L = _lib.lib()
stream = torch.cuda.current_stream().cuda_stream
c_hp = hp.c()
P = flat.numel()


def run(kind, tile_ids, ctas):
    partials = torch.zeros((ctas, P + 16), device="cuda")
    grad, stats = torch.zeros(P, device="cuda"), torch.zeros(8, device="cuda")
    n = tile_ids.numel()
    if kind == "tc":
        _lib.check(L.kin_ppo_grad_tc(flat.data_ptr(), 56, ctypes.byref(c_hp), obs.data_ptr(), act.data_ptr(), old_logp.data_ptr(), adv.data_ptr(),
                                     ret.data_ptr(), sums.data_ptr(), tile_ids.data_ptr(), n, n * 64, partials.data_ptr(), ctas, grad.data_ptr(),
                                     stats.data_ptr(), None, None, 0, 0, None, None, stream))
    else:
        _lib.check(L.kin_ppo_grad(flat.data_ptr(), 56, ctypes.byref(c_hp), obs.data_ptr(), act.data_ptr(), old_logp.data_ptr(), adv.data_ptr(),
                                  ret.data_ptr(), sums.data_ptr(), tile_ids.data_ptr(), n, n * 64, partials.data_ptr(), ctas, grad.data_ptr(),
                                  stats.data_ptr(), None, stream))
    torch.cuda.synchronize()
    return grad, stats


tile_ids = torch.tensor([3, 17, 0, 39, 8, 21, 22, 5, 30, 11, 12, 1, 47, 40], dtype=torch.int32, device="cuda")
idx = (tile_ids.long()[:, None] * 64 + torch.arange(64, device="cuda")[None]).reshape(-1)
lp_out, v_out = torch.full((S,), 123.0, device="cuda"), torch.full((S,), 123.0, device="cuda")
_lib.check(L.kin_ppo_grad_tc(flat.data_ptr(), 56, ctypes.byref(c_hp), obs.data_ptr(), act.data_ptr(), None, None, None, None,
                             tile_ids.data_ptr(), tile_ids.numel(), 0, None, 3, None, None, lp_out.data_ptr(), v_out.data_ptr(), 1, 0, None, None, stream))
torch.cuda.synchronize()
print("forward-only: value max err", float((v_out[idx] - value[idx]).abs().max()), "logp max/mean err",
      float((lp_out[idx] - exact_logp[idx]).abs().max()), float((lp_out[idx] - exact_logp[idx]).abs().mean()), flush=True)
for t in pol.tensors.values():
    t.requires_grad_(True)
loss, ref_stats = _torch_ppo_loss(pol, hp, obs[idx], act[idx], old_logp[idx], adv[idx], ret[idx])
grads = torch.autograd.grad(loss, [pol.tensors[k] for k in ppo.PARAM_ORDER])
for t in pol.tensors.values():
    t.requires_grad_(False)
ref = torch.cat([gk.reshape(-1) for gk in grads])
for kind, ctas in (("fp32", 5), ("tc", 2), ("tc", 7), ("tc", 148)):
    grad, stats = run(kind, tile_ids, ctas)
    off = 0
    rows = []
    for k, gk in zip(ppo.PARAM_ORDER, grads):
        n = gk.numel()
        rows.append(f"{k}:{float((grad[off:off + n] - gk.reshape(-1)).norm() / (gk.norm() + 1e-12)):.2e}")
        off += n
    cos = float(torch.dot(grad, ref) / (grad.norm() * ref.norm() + 1e-30))
    print(kind, ctas, "cos", f"{cos:.6f}", "rel", f"{float((grad - ref).norm() / ref.norm()):.3e}", " ".join(rows))
    print("   stats", [f"{float(x):.5f}" for x in stats[:5]], "ref", {k: round(v, 5) for k, v in ref_stats.items()}, flush=True)

# timing on a realistic minibatch: 131072 samples = 2048 tiles
NT = int(os.environ.get("NT", "2048"))
perm = torch.randperm(S // 64, device="cuda", generator=g)[:NT].to(torch.int32).contiguous()
kinds = ("fp32", "tc")
for kind in kinds:
    for _ in range(3):
        run(kind, perm, 148)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    partials = torch.zeros((148, P + 16), device="cuda")
    grad, stats = torch.zeros(P, device="cuda"), torch.zeros(8, device="cuda")
    n = perm.numel()
    e0.record()
    for _ in range(20):
        if kind == "tc":
            L.kin_ppo_grad_tc(flat.data_ptr(), 56, ctypes.byref(c_hp), obs.data_ptr(), act.data_ptr(), old_logp.data_ptr(), adv.data_ptr(), ret.data_ptr(),
                              sums.data_ptr(), perm.data_ptr(), n, n * 64, partials.data_ptr(), 148, grad.data_ptr(), stats.data_ptr(), None, None, 0, 0, None, None, stream)
        else:
            L.kin_ppo_grad(flat.data_ptr(), 56, ctypes.byref(c_hp), obs.data_ptr(), act.data_ptr(), old_logp.data_ptr(), adv.data_ptr(), ret.data_ptr(),
                           sums.data_ptr(), perm.data_ptr(), n, n * 64, partials.data_ptr(), 148, grad.data_ptr(), stats.data_ptr(), None, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{kind}: {ms * 1e3:.1f} us per {NT * 64}-sample minibatch = {NT * 64 / ms / 1e3:.1f} M samples/s, {NT * 64 * 95.2e3 / ms / 1e9:.1f} TFLOP/s")
