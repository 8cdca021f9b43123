"""Batch-sharded multi-GPU inference (BASELINE.json configs[2]): images are independent (no BatchNorm,
no buffers, the adapter's all-dims fftshift is batch-neutral -- SURVEY.md section 8e), so rank r encodes
images [r*B/n, (r+1)*B/n) with replicated weights and the only exchange step is one all-gather of the
[B/n, out_chans, g, g] embeddings over NCCL / NVLink.  One process per GPU (torchrun env rendezvous,
like the reference's train script, train_ovcos_maskdecoder_edge.py:25-28)."""
from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``total`` images; the first ``total % world`` ranks get one more."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_embeddings(local: torch.Tensor, total: int) -> torch.Tensor:
    """All-gather per-rank embeddings [b_r, C, g, g] into [total, C, g, g] on every rank.  Equal shards use
    one ``all_gather_into_tensor`` (a single NCCL collective); ragged shards pad to the largest shard."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(total, r, world) for r in range(world)]
    counts = [e - s for s, e in sizes]
    biggest = max(counts)
    if local.shape[0] != counts[dist.get_rank()]:
        raise ValueError("local shard size does not match shard_range()")
    if all(c == biggest for c in counts) and dist.get_backend() == "nccl":
        out = local.new_empty((total,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    padded = local.new_zeros((biggest,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], 0)


def encode_sharded(encode: Callable[[torch.Tensor], torch.Tensor], images: torch.Tensor) -> torch.Tensor:
    """Every rank holds (or can index) the same global batch ``images``; each encodes its shard and all ranks
    return the full [B, C, g, g] embedding tensor."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    s, e = shard_range(images.shape[0], rank, world)
    local = encode(images[s:e])
    return gather_embeddings(local, images.shape[0])
