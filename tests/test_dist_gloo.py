"""world_size-2 gloo test (CPU) of the batch-sharding / gather host logic used at N>1.  The encoder
itself has no CPU path, so the per-rank encode function here is the fp32 check path on the tiny config;
what is under test is shard_range / gather_embeddings / encode_sharded."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from camouflaged_vlm_b200.dist import shard_range


def test_shard_range_covers_batch_exactly():
    for total in (1, 5, 32, 33, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (s0, e0), (s1, e1) in zip(spans, spans[1:]):
                assert e0 == s1
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from camouflaged_vlm_b200.check_path import encoder_forward_check
    from camouflaged_vlm_b200.config import PRESETS
    from camouflaged_vlm_b200.dist import encode_sharded
    from camouflaged_vlm_b200.packing import pack_weights
    from camouflaged_vlm_b200.synth import seeded_images, seeded_state_dict
    cfg = PRESETS["tiny"]
    packed = pack_weights(seeded_state_dict(cfg), cfg, "cpu", dtype=torch.float32)
    images = seeded_images(cfg, total, 0)
    out = encode_sharded(lambda x: encoder_forward_check(packed, cfg, x), images)
    if rank == 0:
        full = encoder_forward_check(packed, cfg, images)
        ret["shape"] = tuple(out.shape)
        ret["err"] = float((out - full).abs().max())
    t = torch.tensor([float(out.sum())])
    both = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(both, t)
    if rank == 0:
        ret["same_on_all_ranks"] = bool(all(torch.equal(b, both[0]) for b in both))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [4, 3])     # equal and ragged shards
def test_encode_sharded_world2_gloo(total):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), total, ret), nprocs=world, join=True)
    assert ret["shape"][0] == total
    assert ret["err"] < 1e-5
    assert ret["same_on_all_ranks"]


def _gather_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from camouflaged_vlm_b200.dist import ChunkGatherer
    per_rank, tail = 5, (3, 2, 2)
    full = torch.arange(world * per_rank * 12, dtype=torch.float32).reshape(world * per_rank, *tail) / 7.0
    mine = full[rank * per_rank:(rank + 1) * per_rank]
    ok = True
    for dtype in (torch.float32, torch.bfloat16):
        g = ChunkGatherer(per_rank, tail, "cpu", dtype)
        for s in range(0, per_rank, 2):                      # ragged last chunk (2, 2, 1)
            e = min(per_rank, s + 2)
            g.submit(s, e, mine[s:e])
        out = g.finish()
        ok = ok and tuple(out.shape) == tuple(full.shape) and torch.equal(out, full.to(dtype))
        ok = ok and g.bytes_received == (world - 1) * per_rank * 12 * out.element_size()
    try:
        ChunkGatherer(per_rank, tail, "cpu").submit(4, 6, mine[:2])
        ok = False
    except ValueError:
        pass
    ret[rank] = ok
    dist.destroy_process_group()


def test_chunk_gatherer_world2_gloo():
    """Pipelined per-chunk all-gather (bench.py N>1 path): rank-major layout, ragged chunks, bf16 payload."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gather_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
