"""Multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous.

Inference shards naturally -- each image's heatmap and keypoints depend on that image alone once BN is
folded -- so ranks take contiguous slices of the batch and there is NO data-path collective; only an
optional gather of the (B,K,2) int32 keypoints.  Training is data-parallel with one exchange step: a sum
all-reduce of the gradients (NCCL over NVLink on GPUs, gloo in the CPU tests), averaged over ranks.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of `total` items owned by `rank`; sizes differ by at most one and the
    first `total % world_size` ranks get the extra item."""
    if world_size <= 0 or not 0 <= rank < world_size or total < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(total, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_keypoints(local_yx: torch.Tensor, total: int) -> torch.Tensor:
    """All-gather the per-rank (b_local,K,2) keypoints into (total,K,2) in shard order (host gather, not on the
    data path)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_yx
    world = dist.get_world_size()
    sizes = [shard_range(total, world, r) for r in range(world)]
    pad = max(e - b for b, e in sizes)
    buf = local_yx.new_zeros((pad,) + tuple(local_yx.shape[1:]))
    buf[: local_yx.shape[0]] = local_yx
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[: e - b] for o, (b, e) in zip(out, sizes)], dim=0)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 25 << 20) -> int:
    """Average gradients across ranks with bucketed flat all-reduces.  Returns the number of collectives issued.
    The 21.8 M fp32 gradients (87 MB) go out in ~25 MB buckets so NCCL can overlap them."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    grads: List[torch.Tensor] = [p.grad for p in params if p.grad is not None]
    calls, bucket, size = 0, [], 0

    def flush():
        nonlocal calls, bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for g in bucket:
            n = g.numel()
            g.copy_(flat[off: off + n].view_as(g))
            off += n
        calls += 1
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return calls


def broadcast_model(model: torch.nn.Module, src: int = 0) -> None:
    """Rank `src`'s parameters and buffers overwrite everyone else's (start-of-training sync)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src)
