"""Times the global and the window attention kernel of one side library built by tools/attn_probe.sh (GPU only).
    python tools/attn_probe.py probes/libcvlm_g1.so [batch]
The probes produce garbage results on purpose; only the time matters."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import camouflaged_vlm_b200._lib as _lib

_lib.LIB_PATH = os.path.abspath(sys.argv[1])
import torch

from camouflaged_vlm_b200 import ops


def timeit(fn, iters=5, warm=2):
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
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    M, D, H, hd = B * 4096, 1280, 16, 80
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = (torch.randn(M, 3 * D, generator=g, device="cuda") * 0.7).bfloat16()
    out = {}
    for name, window, s in (("global", 0, 64), ("window", 14, 14)):
        relh = (torch.randn(2 * s - 1, hd, generator=g, device="cuda") * 0.1).bfloat16()
        relw = (torch.randn(2 * s - 1, hd, generator=g, device="cuda") * 0.1).bfloat16()
        out[name] = timeit(lambda: ops.attention(qkv, relh, relw, B, 64, window, H, hd))
    print(f"{os.path.basename(sys.argv[1]):20s} batch {B}: global {out['global']:7.3f} ms   window {out['window']:7.3f} ms", flush=True)


if __name__ == "__main__":
    main()
