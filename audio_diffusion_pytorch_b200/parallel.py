"""Batch data-parallelism of the hot path over the GPUs of one box.

Samples are independent everywhere (GroupNorm per sample, LayerNorm per position, attention
per sample), so inference shards the batch with NO collective; training adds one gradient
all-reduce (NCCL over NVLink) -- either torch DDP around `DiffusionModel`, or the explicit
flat-bucket `allreduce_gradients` below.  One process per GPU (torchrun)."""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist
from torch import Tensor


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced split of n items: first (n % world) ranks get one extra."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(t: Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> Tensor:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def gather_batch(local: Tensor, total: int) -> Tensor:
    """Inverse of shard_batch (ragged shards allowed): every rank gets the full batch."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


@torch.no_grad()
def sample_sharded(model, noise: Tensor, num_steps: int, gather: bool = True, **kwargs) -> Tensor:
    """Each rank runs VSampler on its slice of the batch; no communication inside the loop."""
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_bounds(noise.shape[0], rank, world)
    kw = {k: (v[lo:hi] if torch.is_tensor(v) and v.shape[:1] == noise.shape[:1] else v)
          for k, v in kwargs.items()}
    local = model.sample(noise[lo:hi], num_steps=num_steps, **kw)
    return gather_batch(local, noise.shape[0]) if gather else local


@torch.no_grad()
def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_mb: float = 256.0) -> None:
    """Mean of `.grad` over ranks in flat buckets (each shard's loss is a mean over its own
    samples, so averaging equal-sized shards gives the global-batch gradient)."""
    world = dist.get_world_size()
    bucket: List[Tensor] = []
    nbytes = 0

    def flush():
        nonlocal bucket, nbytes
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        bucket, nbytes = [], 0

    for p in params:
        if p.grad is None:
            continue
        bucket.append(p.grad)
        nbytes += p.grad.numel() * p.grad.element_size()
        if nbytes >= bucket_mb * 2 ** 20:
            flush()
    flush()
