"""Does TMA zero-fill of the padded window rows cost the window-attention kernel time?  Same number of 14x14 windows
(25 per image) with and without padding: token grid 64 (70x70 padded, 16.4 % of the rows out of range) vs 70 (none)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from camouflaged_vlm_b200 import ops


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


B, H, hd = 8, 16, 80
g = torch.Generator(device="cuda").manual_seed(0)
for grid in (56, 64, 70):
    M = B * grid * grid
    qkv = (torch.randn(M, 3 * H * hd, generator=g, device="cuda") * 0.7).bfloat16()
    relh = (torch.randn(27, hd, generator=g, device="cuda") * 0.1).bfloat16()
    relw = (torch.randn(27, hd, generator=g, device="cuda") * 0.1).bfloat16()
    ms = timeit(lambda: ops.attention(qkv, relh, relw, B, grid, 14, H, hd))
    nwin = ((grid + 13) // 14) ** 2
    print(f"grid {grid}: {nwin} windows/img, {ms:.3f} ms, {ms * 1e3 / (B * H * nwin):.3f} us per (window, head) item")
