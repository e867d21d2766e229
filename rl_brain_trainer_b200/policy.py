"""SB3 ``MultiInputPolicy`` actor/critic weights on the device, and the batched ``predict``.

The reference's policies are SB3 PPO checkpoints (``model.zip`` = ``data`` JSON + ``policy.pth`` + optimizer state,
SURVEY F3/F4).  ``PolicyWeights.load`` reads either such a zip (only ``torch.load(weights_only=True)`` is needed, not SB3)
or the ``.npz`` presets exported by ``tests/golden/gen_golden.py``; the state-dict key names are kept so checkpoints
interchange with the reference (``save_npz`` / ``state_dict``).
"""

from __future__ import annotations

import ctypes
import io
import zipfile
from pathlib import Path
from typing import Mapping

import numpy as np
import torch

from . import _lib
from .config import PRESET_DIR

KEYS = {
    "pi_w0": "mlp_extractor.policy_net.0.weight", "pi_b0": "mlp_extractor.policy_net.0.bias",
    "pi_w1": "mlp_extractor.policy_net.2.weight", "pi_b1": "mlp_extractor.policy_net.2.bias",
    "act_w": "action_net.weight", "act_b": "action_net.bias",
    "vf_w0": "mlp_extractor.value_net.0.weight", "vf_b0": "mlp_extractor.value_net.0.bias",
    "vf_w1": "mlp_extractor.value_net.2.weight", "vf_b1": "mlp_extractor.value_net.2.bias",
    "val_w": "value_net.weight", "val_b": "value_net.bias", "log_std": "log_std",
}


class PolicyWeights:
    """fp32 weights resident on one GPU + the ``KinPolicyWeights`` view the kernels take."""

    def __init__(self, state_dict: Mapping[str, np.ndarray | torch.Tensor], device: str | torch.device = "cuda") -> None:
        self.device = torch.device(device)
        self.tensors: dict[str, torch.Tensor] = {}
        for field, key in KEYS.items():
            if key in state_dict:
                self.tensors[field] = torch.as_tensor(np.asarray(state_dict[key]), dtype=torch.float32).contiguous().to(self.device)
        for need in ("pi_w0", "pi_b0", "pi_w1", "pi_b1", "act_w", "act_b"):
            if need not in self.tensors:
                raise KeyError(f"policy state dict lacks {KEYS[need]}")
        self.in_dim = int(self.tensors["pi_w0"].shape[1])
        if tuple(self.tensors["pi_w0"].shape) != (64, self.in_dim) or tuple(self.tensors["pi_w1"].shape) != (64, 64) or \
                tuple(self.tensors["act_w"].shape) != (7, 64):
            raise ValueError("expected the 64-64 tanh MLP of the reference checkpoints (SURVEY F4)")
        self.has_value = all(k in self.tensors for k in ("vf_w0", "vf_b0", "vf_w1", "vf_b1", "val_w", "val_b"))
        self._c = None

    @property
    def c(self):
        if self._c is None:
            w = _lib.c_struct("KinPolicyWeights")()
            w.in_dim = self.in_dim
            w.has_value = int(self.has_value)
            for field, t in self.tensors.items():
                setattr(w, field, t.data_ptr())
            self._c = w
        return self._c

    def state_dict(self) -> dict[str, torch.Tensor]:
        return {KEYS[f]: t for f, t in self.tensors.items()}

    def save_npz(self, path: str | Path) -> None:
        np.savez_compressed(path, **{k: v.detach().cpu().numpy() for k, v in self.state_dict().items()})

    def save_weights_zip(self, path: str | Path, data: dict | None = None, extra: Mapping[str, bytes] | None = None) -> None:
        """Write a WEIGHTS-ONLY zip laid out like an SB3 ``model.zip``: ``policy.pth`` holds the state dict under the reference's key
        names, ``data`` a plain JSON of hyper-parameters.  It is NOT a full SB3 archive -- SB3's ``PPO.load`` also wants the pickled
        observation / action spaces and schedules, which need stable-baselines3 + gymnasium (absent here) to produce.  To take a
        policy trained here into the reference: build ``PPO("MultiInputPolicy", env, **hyper)`` as ``train_workspace_expansion.py:199``
        does and ``model.policy.load_state_dict(torch.load(<policy.pth from this zip>))`` (the keys and shapes are SB3's own);
        ``PPOTrainer.save_checkpoint`` adds ``policy.optimizer.pth`` in torch-Adam layout for ``model.policy.optimizer.load_state_dict``.
        ``PolicyWeights.load`` reads these zips and the reference's own ``model.zip`` checkpoints alike."""
        import json

        buf = io.BytesIO()
        torch.save({k: v.detach().cpu() for k, v in self.state_dict().items()}, buf)
        with zipfile.ZipFile(Path(path), "w") as z:
            z.writestr("policy.pth", buf.getvalue())
            z.writestr("data", json.dumps(data or {"policy_class": "MultiInputPolicy", "n_envs": 1}))
            z.writestr("_stable_baselines3_version", "2.8.0")
            z.writestr("system_info.txt", "written by rl_brain_trainer_b200 (weights-only archive)")
            for name, blob in (extra or {}).items():
                z.writestr(name, blob)

    @classmethod
    def load(cls, path: str | Path, device: str | torch.device = "cuda") -> "PolicyWeights":
        path = Path(path)
        if path.suffix == ".npz":
            with np.load(path) as z:
                return cls({k: z[k] for k in z.files}, device)
        if path.suffix == ".zip":  # SB3 model.zip
            with zipfile.ZipFile(path) as z:
                sd = torch.load(io.BytesIO(z.read("policy.pth")), weights_only=True, map_location="cpu")
            return cls({k: v.numpy() for k, v in sd.items()}, device)
        sd = torch.load(path, weights_only=True, map_location="cpu")
        return cls({k: v.numpy() for k, v in sd.items()}, device)

    @classmethod
    def preset(cls, name: str, device: str | torch.device = "cuda") -> "PolicyWeights":
        """One of the four bundled checkpoints: approach_stage8_11, finisher, randomstart, route_prefix120."""
        return cls.load(PRESET_DIR / "policies" / f"{name}.npz", device)

    # ------------------------------------------------------------------
    def predict(self, obs: torch.Tensor, *, with_value: bool = False) -> tuple[torch.Tensor, torch.Tensor | None]:
        """``model.predict(obs, deterministic=True)`` for a batch: obs [n,in_dim] -> (action [n,7] clipped to +-1, value [n])."""
        o = torch.as_tensor(obs, dtype=torch.float32, device=self.device).reshape(-1, self.in_dim).contiguous()
        n = o.shape[0]
        action = torch.empty((n, 7), dtype=torch.float32, device=self.device)
        value = torch.empty(n, dtype=torch.float32, device=self.device) if (with_value and self.has_value) else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().kin_policy_forward(ctypes.byref(self.c), o.data_ptr(), action.data_ptr(),
                                                     None if value is None else value.data_ptr(), n,
                                                     torch.cuda.current_stream().cuda_stream))
        return action, value
