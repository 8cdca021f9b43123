"""Attention timing vs head_dim (64: one TMA row per token and operand; 80: two rows, the second only 32 bytes).
If the kernels are bound by the TMA row rate, head_dim 64 is ~2x faster although the softmax work is identical."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from camouflaged_vlm_b200 import ops


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


B, H = 8, 16
g = torch.Generator(device="cuda").manual_seed(0)
for hd in (64, 80):
    D = H * hd
    qkv = (torch.randn(B * 4096, 3 * D, generator=g, device="cuda") * 0.7).bfloat16()
    for name, window, s in (("global", 0, 64), ("window", 14, 14)):
        relh = (torch.randn(2 * s - 1, hd, generator=g, device="cuda") * 0.1).bfloat16()
        relw = (torch.randn(2 * s - 1, hd, generator=g, device="cuda") * 0.1).bfloat16()
        ms = timeit(lambda: ops.attention(qkv, relh, relw, B, 64, window, H, hd), iters=5, warm=2)
        print(f"head_dim {hd} {name}: {ms:.3f} ms")
