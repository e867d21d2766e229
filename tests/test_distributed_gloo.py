"""world_size-2 gloo tests (CPU) of the multi-rank host logic: sharding, gradient / statistics all-reduce, curriculum agreement."""

from __future__ import annotations

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rl_brain_trainer_b200 import distributed as kd

    # 1. episode sharding: every unit owned exactly once, contiguous
    n_total = 1001
    sl = kd.shard_slice(n_total, rank, world)
    owned = torch.zeros(n_total)
    owned[sl] = 1
    kd.allreduce_sum_(owned)
    assert torch.equal(owned, torch.ones(n_total))
    # 2. gradient all-reduce: each rank holds the sum over ITS samples of per-sample grads / global batch
    g = torch.Generator().manual_seed(0)
    per_sample = torch.randn(64, 16143, generator=g)              # same on both ranks
    mine = per_sample[rank::world]
    grad = mine.sum(0) / 64.0
    kd.allreduce_sum_(grad)
    assert torch.allclose(grad, per_sample.mean(0), atol=1e-6)
    # 3. eval statistics: success rate over the union of the shards
    success = (torch.arange(n_total) % 3 == 0)[sl]
    pos = torch.full((success.numel(),), 0.002 * (rank + 1))
    stats = kd.reduce_eval_stats(success, pos, pos * 2, env_steps=int(success.numel()) * 164)
    assert stats["episodes"] == n_total and abs(stats["success_rate"] - float((torch.arange(n_total) % 3 == 0).double().mean())) < 1e-12
    assert stats["env_steps"] == n_total * 164
    # 4. curriculum: ranks see different local outcomes but promote together
    tr = kd.CurriculumTracker(n_stages=4, success_rate_threshold=0.68, window_episodes=32, min_episodes_per_stage=64, stage_index=0)
    promoted = []
    for it in range(4):
        local_succ = 60 if rank == 0 else 20          # 80 / 128 = 0.625 < 0.68 globally; rank 0 alone would promote
        promoted.append(tr.record(float(local_succ + 10 * it), 64.0))
    np.save(os.path.join(out_dir, f"promoted_{rank}.npy"), np.array(promoted))
    mx = torch.tensor([float(rank)])
    kd.allreduce_max_(mx)
    assert float(mx) == world - 1
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "promoted_0.npy"), np.load(tmp_path / "promoted_1.npy")
    assert np.array_equal(a, b)            # identical decisions on both ranks
    assert a.tolist() == [False, True, False, True] or a.sum() >= 1


def test_single_process_helpers():
    from rl_brain_trainer_b200 import distributed as kd

    assert kd.world() == (0, 1)
    parts = [kd.shard_slice(10, r, 4) for r in range(4)]
    assert [p.stop - p.start for p in parts] == [3, 3, 2, 2] and parts[0].start == 0 and parts[-1].stop == 10
    t = torch.ones(3)
    assert kd.allreduce_sum_(t) is t
    tr = kd.CurriculumTracker(3, 0.8, 20, 30)
    assert not tr.record(10, 20)            # below min_episodes_per_stage
    assert tr.record(19, 20) and tr.stage_index == 1
