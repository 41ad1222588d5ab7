"""Batch sharding of the head + post-processing path over the GPUs of one box.

Every image's normalise / similarity / decode / threshold / NMS is independent of every other
image (model/yolo_clip.py:177-206 has no cross-batch op; inference/detector.py:163-223 is per
image), so the path shards by image with **no collective on the data path**: one process per
GPU (``torchrun``), rank r owns a contiguous slice of the batch, the vocabulary is replicated.
``torch.distributed`` is used only around the path:

* ``broadcast_vocabulary``  once per vocabulary (rank 0 -> all),
* ``gather_detections``     optional, to hand the per-image results to rank 0,
* ``max_over_ranks``        the timing reduction of ``bench.py``.

The functions are backend-agnostic (``nccl`` on the B200s, ``gloo`` in the CPU tests); the
per-rank compute is whatever the caller passes (``Detector.predict`` in production).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice ``[lo, hi)`` of ``total`` images for ``rank``; the first ``total % world``
    ranks take one extra image, so slices are balanced, ordered by rank, and cover the batch
    exactly (an empty slice when ``total < world``)."""
    if world <= 0 or not 0 <= rank < world or total < 0:
        raise ValueError(f"ovdet: bad shard request total={total} rank={rank} world={world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total: int, world: int) -> List[int]:
    return [hi - lo for lo, hi in (shard_range(total, r, world) for r in range(world))]


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_batch(tensors: Sequence[torch.Tensor], group=None) -> List[torch.Tensor]:
    """This rank's slice (a view, dim 0) of each ``[B, ...]`` tensor."""
    rank, world = _world(group)
    lo, hi = shard_range(tensors[0].shape[0], rank, world)
    return [t[lo:hi] for t in tensors]


def broadcast_vocabulary(text: Optional[torch.Tensor], num_classes: int, embed_dim: int, device,
                         src: int = 0, group=None) -> torch.Tensor:
    """Replicate the ``[C, D]`` fp32 vocabulary from ``src`` to every rank (once per vocabulary,
    outside any timed region; C x 512 x 4 B = 2.5 MB at C = 1203)."""
    rank, world = _world(group)
    if rank == src:
        buf = text.to(device=device, dtype=torch.float32).contiguous()
        if buf.shape != (num_classes, embed_dim):
            raise ValueError("ovdet: vocabulary shape does not match (num_classes, embed_dim)")
    else:
        buf = torch.empty(num_classes, embed_dim, device=device, dtype=torch.float32)
    if world > 1:
        dist.broadcast(buf, src=src, group=group)
    return buf


def max_over_ranks(value: float, device, group=None) -> float:
    """Elapsed time of the slowest rank (the denominator of whole-job throughput)."""
    _, world = _world(group)
    if world == 1:
        return float(value)
    t = torch.tensor([value], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_detections(local: Dict[str, torch.Tensor], total: int, dst: int = 0,
                      group=None) -> Optional[Dict[str, torch.Tensor]]:
    """Collect per-image result tensors (dim 0 = this rank's images, any trailing shape, same
    keys on every rank) on ``dst`` in global image order.  Shards may be ragged: every rank pads
    to the largest shard, ``dst`` trims.  Returns the gathered dict on ``dst``, None elsewhere."""
    rank, world = _world(group)
    if world == 1:
        return {k: v for k, v in local.items()}
    sizes = shard_sizes(total, world)
    cap = max(sizes)
    out: Dict[str, torch.Tensor] = {}
    for key in sorted(local):
        v = local[key]
        if v.shape[0] != sizes[rank]:
            raise ValueError(f"ovdet: `{key}` has {v.shape[0]} images, shard holds {sizes[rank]}")
        padded = torch.zeros((cap,) + tuple(v.shape[1:]), device=v.device, dtype=v.dtype)
        padded[: v.shape[0]] = v
        parts = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
        dist.gather(padded, parts, dst=dst, group=group)
        if rank == dst:
            out[key] = torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
    return out if rank == dst else None


def run_sharded(compute: Callable[..., Dict[str, torch.Tensor]], batch_tensors: Sequence[torch.Tensor],
                gather: bool = True, group=None) -> Optional[Dict[str, torch.Tensor]]:
    """Slice the global batch, run ``compute(*local_slices)`` on this rank (no communication
    inside), optionally gather the per-image results on rank 0."""
    total = batch_tensors[0].shape[0]
    local = compute(*shard_batch(batch_tensors, group))
    if not gather:
        return local
    return gather_detections(local, total, group=group)
