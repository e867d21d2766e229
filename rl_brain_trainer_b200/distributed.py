"""Multi-GPU plumbing: envs / episodes shard by index, collectives only where the path has a real exchange.

* evaluation: no data-path collective; one all-reduce (sum) of a small statistics vector at the end.
* training: one all-reduce (sum) of the flat gradient (+ the statistics) per minibatch; every rank then applies the
  identical Adam step.  Curriculum promotion uses an all-reduced (successes, episodes) pair so all ranks switch stage together.

All helpers take plain tensors and an optional process group, so they run under NCCL (one rank per GPU) and under gloo on
CPU tensors (tests/test_distributed_gloo.py, world_size 2).

``PeerGradExchange`` is the NVLink peer-memory form of the training all-reduce (``csrc/kin_peer.cu``): the reduction kernel pushes
the rank's gradient straight into every peer's receive buffer and a gather kernel sums the slots in rank order -- two small
kernels per minibatch instead of a latency-bound NCCL call; ``torch.distributed`` only carries the one-off IPC handle exchange.
"""

from __future__ import annotations

from typing import Any

import torch
import torch.distributed as dist


def world(group: Any = None) -> tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_slice(n_total: int, rank: int, world_size: int) -> slice:
    """Contiguous shard of ``n_total`` independent units (episodes, envs, route replicas) owned by ``rank``.

    The first ``n_total % world_size`` ranks get one extra unit, so every unit is owned exactly once.
    """
    base, extra = divmod(int(n_total), int(world_size))
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def allreduce_sum_(t: torch.Tensor, group: Any = None) -> torch.Tensor:
    """In-place sum over ranks (no-op for a single process)."""
    if world(group)[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allreduce_max_(t: torch.Tensor, group: Any = None) -> torch.Tensor:
    if world(group)[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


def reduce_eval_stats(success: torch.Tensor, final_pos: torch.Tensor, final_ori: torch.Tensor, env_steps: torch.Tensor | int,
                      group: Any = None) -> dict[str, float]:
    """Success rate / mean errors / env-steps over all ranks from per-rank episode rows (one small all-reduce)."""
    v = torch.stack([success.double().sum(), torch.as_tensor(float(success.numel()), dtype=torch.float64, device=success.device),
                     final_pos.double().sum(), final_ori.double().sum(),
                     torch.as_tensor(env_steps, dtype=torch.float64, device=success.device).reshape(())])
    allreduce_sum_(v, group)
    n = max(float(v[1]), 1.0)
    return {"episodes": float(v[1]), "success_rate": float(v[0]) / n, "mean_final_position_error": float(v[2]) / n,
            "mean_final_orientation_error": float(v[3]) / n, "env_steps": float(v[4])}


class PeerGradExchange:
    """Per-minibatch gradient all-reduce (sum) over CUDA-IPC peer buffers of the GPUs of one node (``kin_peer_*`` in the C ABI).

    ``push(partials, n_cta, global_batch)`` reduces this rank's per-CTA rows and stores them into every rank's buffer;
    ``gather(grad, stats)`` waits on the device for all ranks and writes the rank-ordered sum (bitwise identical everywhere).
    Both only enqueue kernels on the current stream.  ``check()`` raises if a peer never arrived (device-side timeout).
    """

    def __init__(self, n_params: int, device: torch.device, group: Any = None) -> None:
        import ctypes

        from . import _lib

        self._L, self._check = _lib.lib(), _lib.check
        self.rank, self.world = world(group)
        if self.world > 8:
            raise _lib.KinError("PeerGradExchange: at most 8 ranks (the GPUs of one NVSwitch node)")
        self.P, self.device = int(n_params), torch.device(device)
        self.epoch = 0
        with torch.cuda.device(self.device):
            own = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(_lib.define("KIN_PEER_HANDLE_BYTES"))
            self._check(self._L.kin_peer_buffer_create(self.P, self.world, ctypes.byref(own), handle))
            self._own = own.value
            handles: list[Any] = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self._bufs = (ctypes.c_void_p * self.world)()
            self._opened: list[int] = []
            for r in range(self.world):
                if r == self.rank:
                    self._bufs[r] = self._own
                else:
                    ptr = ctypes.c_void_p()
                    self._check(self._L.kin_peer_buffer_open(handles[r], ctypes.byref(ptr)))
                    self._bufs[r] = ptr.value
                    self._opened.append(ptr.value)
            self.timed_out = torch.zeros(1, dtype=torch.int32, device=self.device)
            torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=group)        # every buffer exists and is zeroed before the first push can reach it

    def push(self, partials: torch.Tensor, n_cta: int, global_batch: int) -> None:
        self.epoch += 1
        self._check(self._L.kin_peer_grad_push(partials.data_ptr(), int(n_cta), self.P, int(global_batch), self._bufs, self.rank, self.world,
                                               self.epoch, torch.cuda.current_stream(self.device).cuda_stream))

    def gather(self, grad: torch.Tensor, stats: torch.Tensor | None) -> None:
        self._check(self._L.kin_peer_grad_gather(self._own, self.P, self.world, self.epoch, grad.data_ptr(),
                                                 None if stats is None else stats.data_ptr(), self.timed_out.data_ptr(),
                                                 torch.cuda.current_stream(self.device).cuda_stream))

    def check(self) -> None:
        if int(self.timed_out.item()):
            from . import _lib

            raise _lib.KinError("PeerGradExchange: a peer rank never delivered its gradient (device-side wait timed out)")

    def close(self) -> None:
        """Unmap the peers' buffers and free this rank's (after the stream drained).  Every rank must have finished its last
        gather before any rank closes: call it at the same point of the program on all ranks (``PPOTrainer.close`` does)."""
        if getattr(self, "_own", None) is None:
            return
        torch.cuda.synchronize(self.device)
        for ptr in self._opened:
            self._L.kin_peer_buffer_close(ptr)
        self._L.kin_peer_buffer_destroy(self._own)
        self._own, self._opened = None, []

    def __del__(self) -> None:  # pragma: no cover - interpreter shutdown order is not ours
        try:
            self.close()
        except Exception:
            pass


class CurriculumTracker:
    """Windowed success-rate promotion (``PointCurriculumTracker`` / ``PointCurriculumCallback``, curriculum.py:104-154,
    training/callbacks.py:54-92), fed with per-iteration (successes, episodes) counts that are summed over ranks first.

    The reference keeps a deque of the last ``window_episodes`` episode outcomes; with tens of thousands of envs finishing
    together the window is filled many times per iteration, so the rate of the latest batch of finished episodes is used
    when that batch alone is at least a window long, otherwise batches are accumulated until it is.
    """

    def __init__(self, n_stages: int, success_rate_threshold: float, window_episodes: int, min_episodes_per_stage: int, stage_index: int = 0) -> None:
        self.n_stages, self.threshold = int(n_stages), float(success_rate_threshold)
        self.window, self.min_episodes = int(window_episodes), int(min_episodes_per_stage)
        self.stage_index = int(stage_index)
        self.stage_episode_count = 0
        self._succ = 0.0
        self._eps = 0.0
        self.history: list[dict[str, float | int]] = []

    def record(self, successes: torch.Tensor | float, episodes: torch.Tensor | float, group: Any = None) -> bool:
        v = torch.as_tensor([float(successes), float(episodes)], dtype=torch.float64)
        if world(group)[1] > 1:
            dev = successes.device if isinstance(successes, torch.Tensor) else None
            backend = dist.get_backend(group)
            v = v.to(dev) if (backend == "nccl" and dev is not None) else v
            allreduce_sum_(v, group)
        s, e = float(v[0]), float(v[1])
        if e <= 0:
            return False
        if e >= self.window:
            self._succ, self._eps = s, e
        else:
            self._succ += s
            self._eps += e
        self.stage_episode_count += int(e)
        if self.stage_index >= self.n_stages - 1 or self.stage_episode_count < self.min_episodes or self._eps < self.window:
            return False
        rate = self._succ / self._eps
        if rate < self.threshold:
            if self._eps >= 4 * self.window:
                self._succ, self._eps = 0.0, 0.0
            return False
        self.history.append({"from_stage_index": self.stage_index, "to_stage_index": self.stage_index + 1, "trigger_success_rate": rate})
        self.stage_index += 1
        self.stage_episode_count = 0
        self._succ, self._eps = 0.0, 0.0
        return True
