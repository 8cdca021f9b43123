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


class ChunkGatherer:
    """Pipelined embedding all-gather ("overlap with next micro-batch on a side stream").

    A rank produces its ``per_rank`` embeddings chunk by chunk; ``submit(s, e, chunk)`` is called right after the
    encoder of images [s, e) has been enqueued and starts the all-gather of that chunk on a side stream, so the
    transfer runs underneath the encoder of the next chunk instead of after the whole batch.  Every rank ends up with
    ``out`` = [world * per_rank, C, g, g] in rank-major order (rank r's image i at r * per_rank + i), the layout one
    ``all_gather_into_tensor`` of the whole shard would give.  ``dtype=torch.bfloat16`` halves the bytes on the wire
    (the embedding is rounded to bf16: ~1.1e-3 rel-L2 on top of the encoder's own error); ``finish()`` makes the
    current stream wait for every gather and returns ``out``.  On CPU tensors (gloo, tests) everything is synchronous."""

    def __init__(self, per_rank: int, tail, device, dtype=torch.float32):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.per_rank, self.dtype = per_rank, dtype
        self.out = torch.empty((self.world * per_rank,) + tuple(tail), device=device, dtype=dtype)
        self._by_rank = self.out.view((self.world, per_rank) + tuple(tail))
        self.stream = torch.cuda.Stream(device=device) if torch.device(device).type == "cuda" else None
        self.bytes_received = 0

    def submit(self, s: int, e: int, chunk: torch.Tensor) -> None:
        if chunk.shape[0] != e - s or not (0 <= s < e <= self.per_rank):
            raise ValueError(f"chunk of {chunk.shape[0]} images does not cover [{s}, {e}) of {self.per_rank}")
        if self.world == 1:
            self._by_rank[0, s:e].copy_(chunk)
            return
        views = [self._by_rank[r, s:e] for r in range(self.world)]         # contiguous slices of `out`
        self.bytes_received += (self.world - 1) * views[0].numel() * views[0].element_size()
        if self.stream is None:
            dist.all_gather(views, chunk.to(self.dtype).contiguous())
            return
        produced = torch.cuda.Event()
        produced.record(torch.cuda.current_stream(chunk.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(produced)
            src = chunk.to(self.dtype).contiguous()
            dist.all_gather(views, src)
        chunk.record_stream(self.stream)

    def finish(self) -> torch.Tensor:
        if self.stream is not None:
            torch.cuda.current_stream(self.out.device).wait_stream(self.stream)
        return self.out
