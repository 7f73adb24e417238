"""Env-level data parallelism across GPUs: contiguous env-index ranges per rank, no data-path
collective; an optional gather of the per-step outputs to one rank.  Works with any torch.distributed backend
(NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def local_env_range(num_envs_total: int, rank: int, world_size: int):
    """Rank r owns [r*n, (r+1)*n) with n = ceil(N / world); the last rank may own fewer."""
    n = -(-num_envs_total // world_size)
    lo = min(rank * n, num_envs_total)
    return lo, min(lo + n, num_envs_total)


def env_seeds(seed0: int, rank: int, world_size: int, num_envs_total: int):
    """Seeds `seed0 + global env index`, so results do not depend on the world size."""
    lo, hi = local_env_range(num_envs_total, rank, world_size)
    return list(range(seed0 + lo, seed0 + hi))


def gather_step_outputs(obs: dict, reward: torch.Tensor, dst: int = 0):
    """all_gather `observation | achieved_goal | desired_goal | reward` rows (equal local sizes) and return the
    [world*n, dim] tensor on every rank (rank `dst` is the consumer)."""
    if isinstance(obs, dict):
        packed = torch.cat([obs["observation"], obs["achieved_goal"], obs["desired_goal"], reward[:, None]], dim=1).contiguous()
    else:   # flat observation (Adroit)
        packed = torch.cat([obs, reward[:, None]], dim=1).contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return packed
    out = torch.empty((dist.get_world_size() * packed.shape[0], packed.shape[1]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed)
    return out


def mixed_batch_assignment(env_ids, world_size: int):
    """BASELINE config 5 style mixed batches (heterogeneous models): whole ranks per env id, split as evenly
    as the world size allows (e.g. 4 GPUs, 2 ids -> ranks 0, 1 run the first id and ranks 2, 3 the second).  Returns the
    env id of every rank."""
    k = len(env_ids)
    if world_size < k:
        raise ValueError("a mixed batch needs at least one rank per env id")
    return [env_ids[min(r * k // world_size, k - 1)] for r in range(world_size)]


def gather_mixed_outputs(obs: torch.Tensor, reward: torch.Tensor, width: int):
    """all_gather of flat observations whose width differs between ranks (one model per rank): rows are zero-padded to
    `width` (the widest observation of the batch) and the reward is appended; [world * n, width + 1] on every rank."""
    n, d = obs.shape
    packed = torch.zeros((n, width + 1), dtype=obs.dtype, device=obs.device)
    packed[:, :d] = obs
    packed[:, width] = reward
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return packed
    out = torch.empty((dist.get_world_size() * n, width + 1), dtype=obs.dtype, device=obs.device)
    dist.all_gather_into_tensor(out, packed)
    return out


class PackedGather:
    """The optional all-gather of the per-step results to every rank (rank 0 is the consumer), on the PACKED rows of
    the step kernel (`b200sim_set_packed`: obs | achieved | desired | reward | success | terminated | truncated) -- one NCCL
    `all_gather_into_tensor` per step, issued on a side stream so that it overlaps the next step's kernel instead of sitting in
    front of it.  Two destination buffers alternate: the consumer reads `latest()` while the next gather fills the other one."""

    def __init__(self, n_local: int, width: int, device, world_size=None):
        self.world = world_size or (dist.get_world_size() if dist.is_initialized() else 1)
        self.rows = self.world * n_local
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self.bufs = [torch.empty((self.rows, width), dtype=torch.float32, device=self.device) for _ in range(2)]
        self.k, self.done = 0, None

    def launch(self, packed: torch.Tensor):
        """Start the gather of this step's packed rows (produced on the current stream); returns the destination buffer."""
        dst = self.bufs[self.k]
        self.k ^= 1
        if self.stream is None:   # CPU tensors (gloo tests): synchronous
            if self.world > 1:
                dist.all_gather_into_tensor(dst, packed.contiguous())
            else:
                dst.copy_(packed)
            self.dst = dst
            return dst
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        packed.record_stream(self.stream)   # the caching allocator must not hand the rows out again before the gather has read them
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            if self.world > 1:
                dist.all_gather_into_tensor(dst, packed)
            else:
                dst.copy_(packed, non_blocking=True)
            self.done = torch.cuda.Event()
            self.done.record(self.stream)
        self.dst = dst
        return dst

    def wait(self):
        """Block the host until the last launched gather has landed; returns its buffer ([world * n, W])."""
        if self.done is not None:
            self.done.synchronize()
        return getattr(self, "dst", None)
