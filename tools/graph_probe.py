"""Does capturing the forward in a CUDA graph help at small batch?  (launch gaps between 243 dependent kernels)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from camouflaged_vlm_b200 import PRESETS, build_encoder

cfg = PRESETS["vit_h"]
dev = torch.device("cuda")
model = build_encoder(cfg).eval().to(dev)
gen = torch.Generator(device=dev).manual_seed(0)
with torch.no_grad():
    for n, p in model.named_parameters():
        p.copy_(torch.randn(p.shape, generator=gen, device=dev) * 0.02)
        if n.endswith(("norm1.weight", "norm2.weight", "neck.1.weight", "neck.3.weight")):
            p.add_(1.0)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for B in (1, 2, 8):
    x = torch.randn(B, 3, 1024, 1024, generator=gen, device=dev)
    with torch.no_grad():
        eager = timeit(lambda: model(x, interm=True))
        want = model(x, interm=True)[0].clone()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            model(x, interm=True)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            out = model(x, interm=True)
        graphed = timeit(g.replay)
        g.replay()
        torch.cuda.synchronize()
        same = torch.equal(out[0], want)
    print(f"batch {B}: eager {eager:.3f} ms, graph {graphed:.3f} ms, identical {same}")
