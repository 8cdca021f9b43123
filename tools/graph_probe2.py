"""Batch-1 latency of the mask decoder and the CLIP tower, eager vs CUDA-graph replay (camouflaged_vlm_b200.graphed.GraphedCall)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from camouflaged_vlm_b200 import mask_decoder as md                 # noqa: E402
from camouflaged_vlm_b200.graphed import GraphedCall                # noqa: E402
from tower_bench import build_tower                                 # noqa: E402


def timeit(fn, n=30, w=5):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


gen = torch.Generator().manual_seed(0)
dec = md.build_mask_decoder(seed=0).cuda()
emb = torch.randn(1, 256, 64, 64, generator=gen).cuda()
pe = torch.randn(1, 256, 64, 64, generator=gen).cuda()
sparse = torch.randn(1, 2, 256, generator=gen).cuda()
dense = torch.zeros(1, 256, 64, 64).cuda()
dfn = lambda e, s: md.postprocess_masks(dec(e, [], pe, s, dense, False)[0], 1024, 1024, 1024)
with torch.no_grad():
    print(f"decoder batch 1: eager {timeit(lambda: dfn(emb, sparse)):.3f} ms", end="")
g = GraphedCall(dfn, emb, sparse)
print(f", graph {timeit(lambda: g(emb, sparse)):.3f} ms")
tw = build_tower()
img = torch.randn(1, 3, 336, 336, generator=gen).cuda()
al = torch.ones(1, 1, 336, 336).cuda()
shared = (torch.randn(2, 1024, generator=gen) * 0.02).cuda()
deep = [(torch.randn(2, 1024, generator=gen) * 0.02).cuda() for _ in range(8)]
tfn = lambda x, a: tw(x, a, shared, deep)
with torch.no_grad():
    print(f"tower batch 1: eager {timeit(lambda: tfn(img, al)):.3f} ms", end="")
tg = GraphedCall(tfn, img, al)
print(f", graph {timeit(lambda: tg(img, al)):.3f} ms")
