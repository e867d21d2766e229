"""Two ranks on two GPUs of one node (skipped on a single-GPU box): the NVLink peer-memory gradient exchange in the gradient kernel's
tail -- the flag-in-data form, with and without the fused Adam step -- against the two-kernel form and the NCCL all-reduce."""

from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return int(s.getsockname()[1])


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs of one node")
def test_peer_gradient_exchange_two_ranks():
    env = dict(os.environ)
    env.pop("RANK", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tools" / "peer_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = next(l for l in reversed(r.stdout.splitlines()) if l.startswith("{"))
    out = json.loads(line)
    assert out["world"] == 2
    assert out["peer_params_bitwise_identical_across_ranks"] is True        # every rank holds the same parameters, bit for bit
    assert out["fused_tail_bitwise_equals_two_kernel"] is True              # in-kernel exchange = push + gather kernels (same summation orders)
    assert out["peer_vs_nccl_rel_diff"] < 1e-5 and out["one_launch_vs_two_kernel_rel_diff"] < 1e-6
    for k in ("nccl", "peer", "peer_separate_adam", "peer_two_kernel"):
        assert abs(out[k]["value_loss"] - out["nccl"]["value_loss"]) < 1e-5 and abs(out[k]["grad_norm"] - out["nccl"]["grad_norm"]) < 1e-4
