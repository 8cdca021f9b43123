"""Time single kernels at ViT-H shapes through the op-level C ABI (CUDA events, L2 flushed by working on
buffers larger than L2).  GPU only.  Usage: python tools/op_bench.py [batch]"""
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


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    M, D, H, hd = B * 4096, 1280, 16, 80
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device="cuda") * sc)
    a = rn(M, D).bfloat16()
    res = rn(M, D)
    only_attn = len(sys.argv) > 2 and sys.argv[2] == "attn"
    print(f"batch {B}: M = {M}")
    for name, n, k, kw in () if only_attn else (("qkv", 3 * D, D, dict(out_dtype=torch.bfloat16)),
                           ("proj", D, D, dict(residual=res)),
                           ("lin1", 4 * D, D, dict(out_dtype=torch.bfloat16, gelu=True)),
                           ("lin2", D, 4 * D, dict(residual=res))):
        x = rn(M, k).bfloat16() if k != D else a
        w = rn(n, k, sc=0.02).bfloat16()
        bias = rn(n, sc=0.02)
        out = torch.empty(M, n, device="cuda", dtype=kw.get("out_dtype", torch.float32))
        kw2 = {kk: vv for kk, vv in kw.items() if kk != "out_dtype"}
        ms = timeit(lambda: ops.gemm_bf16(x, w, bias=bias, out=out, **kw2))
        print(f"  gemm {name:5s} n={n:5d} k={k:5d}: {ms:7.3f} ms  {2.0 * M * n * k / ms / 1e9:8.1f} TFLOP/s")
        del x, w, out
    qkv = rn(M, 3 * D, sc=0.7).bfloat16()
    for name, window, s in (("global", 0, 64), ("window", 14, 14)):
        relh = rn(2 * s - 1, hd, sc=0.1).bfloat16()
        relw = rn(2 * s - 1, hd, sc=0.1).bfloat16()
        ms = timeit(lambda: ops.attention(qkv, relh, relw, B, 64, window, H, hd), iters=5, warm=2)
        keys = 4096 if window == 0 else 196
        fl = B * H * (4.0 * 4096 * keys * hd + 2.0 * 4096 * 2 * s * hd)
        print(f"  attention {name}: {ms:7.3f} ms  {fl / ms / 1e9:8.1f} algorithmic TFLOP/s")
    x = rn(M, D)
    w, b_ = rn(D), rn(D)
    ms = timeit(lambda: ops.layernorm(x, w, b_, 1e-6))
    print(f"  layernorm: {ms:7.3f} ms  {M * D * 6 / ms / 1e6:8.1f} GB/s")


if __name__ == "__main__":
    main()
