"""CPU tests of the host-side PPO helpers (no GPU, no library calls)."""

from __future__ import annotations

import numpy as np
import torch

from rl_brain_trainer_b200 import ppo


def _encode_images(rows: torch.Tensor) -> torch.Tensor:
    """[G,128,64] float -> [G,16384] uint8 SWIZZLE_128B bf16 images (the layout kin_ppo_collect writes; csrc/kin_umma.cuh)."""
    g = rows.shape[0]
    raw = rows.bfloat16().view(torch.int16).reshape(g, 128, 8, 8)           # row, logical 16-byte chunk, 8 bf16
    out = torch.zeros_like(raw)
    for r in range(128):
        for c in range(8):
            out[:, r, c ^ (r & 7)] = raw[:, r, c]
    return out.reshape(g, 128 * 64).view(torch.uint8)


def test_decode_obs_images_inverts_the_swizzle():
    gen = torch.Generator().manual_seed(0)
    rows = torch.randn((3, 128, 64), generator=gen)
    rows[..., 56] = 1.0
    rows[..., 57:] = 0.0
    img = _encode_images(rows)
    assert img.shape == (3, 16384)
    back = ppo.decode_obs_images(img)
    assert back.shape == (3, 128, 64)
    assert torch.equal(back, rows.bfloat16().float())
    # leading dimensions are kept ([T, tiles, bytes] in the trainer)
    assert ppo.decode_obs_images(img.reshape(1, 3, 16384)).shape == (1, 3, 128, 64)
    # the product-side encoder (fp32 observations -> the images kin_ppo_collect would have written) agrees with the test's scalar one
    assert torch.equal(ppo.encode_obs_images(rows[..., :56].reshape(-1, 56)), img)


def test_numpy_gae_matches_a_scalar_loop():
    rng = np.random.default_rng(1)
    T, n = 9, 4
    rew, val = rng.normal(size=(T, n)), rng.normal(size=(T, n))
    starts = (rng.random((T, n)) < 0.3).astype(np.uint8)
    last_val, last_done = rng.normal(size=n), rng.random(n) < 0.5
    adv, ret = ppo.numpy_gae(rew, val, starts, last_val, last_done, 0.97, 0.9)
    for e in range(n):
        gae = 0.0
        for t in reversed(range(T)):
            nnt = 1.0 - (float(last_done[e]) if t == T - 1 else float(starts[t + 1, e]))
            nv = last_val[e] if t == T - 1 else val[t + 1, e]
            delta = rew[t, e] + 0.97 * nv * nnt - val[t, e]
            gae = delta + 0.97 * 0.9 * nnt * gae
            assert abs(adv[t, e] - gae) < 1e-12 and abs(ret[t, e] - (gae + val[t, e])) < 1e-12


def test_hyper_from_config_and_param_order():
    hp = ppo.PPOHyper.from_config({"learning_rate": 4e-6, "n_steps": 1024, "batch_size": 256, "n_epochs": 8, "gamma": 0.98,
                                   "gae_lambda": 0.95, "clip_range": 0.2, "ent_coef": 0.0, "policy": "MultiInputPolicy"})
    assert hp.learning_rate == 4e-6 and hp.n_steps == 1024 and hp.n_epochs == 8 and hp.vf_coef == 0.5 and hp.max_grad_norm == 0.5
    assert ppo.PARAM_ORDER[0] == "pi_w0" and ppo.PARAM_ORDER[-1] == "log_std" and len(ppo.PARAM_ORDER) == 13
    # 56-input policy: 2 x (64*56 + 64 + 4096 + 64) + 7*64 + 7 + 64 + 1 + 7 parameters (16 143)
    assert 2 * (64 * 56 + 64 + 4096 + 64) + 7 * 64 + 7 + 64 + 1 + 7 == 16143
