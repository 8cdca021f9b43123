"""Diagnose the tcgen05 GEMM on the smallest shapes: prints the error structure so that a wrong
descriptor / swizzle / K-advance can be told apart from one run.  GPU only."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from camouflaged_vlm_b200 import ops


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def probe(m, n, k, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(m, k, generator=g).cuda().bfloat16()
    b = (torch.randn(n, k, generator=g) * 0.1).cuda().bfloat16()
    want = a.float() @ b.float().t()
    got = ops.gemm_bf16(a, b, n=n)
    torch.cuda.synchronize()
    e = rel(got, want)
    print(f"[probe] m={m} n={n} k={k}: rel err {e:.3e}  finite={bool(torch.isfinite(got).all())}")
    if e < 1e-4:
        return True
    print("  got [0,:8] ", [round(v, 4) for v in got[0, :8].tolist()])
    print("  want[0,:8] ", [round(v, 4) for v in want[0, :8].tolist()])
    # hypothesis: only some 16-wide K chunks were accumulated
    for lo in range(0, min(k, 64), 16):
        part = a[:, lo:lo + 16].float() @ b[:, lo:lo + 16].float().t()
        print(f"  vs partial K[{lo}:{lo+16}] only: {rel(got, part):.3e}")
    # hypothesis: rows or columns permuted
    gm, wm = got[:128, :256], want[:128, :256]
    d = torch.cdist(gm.double(), wm.double())
    rp = d.argmin(1)
    print("  best matching want-row for got rows 0..15:", rp[:16].tolist(), " min dist", float(d.min(1).values.max()))
    d = torch.cdist(gm.t().double().contiguous(), wm.t().double().contiguous())
    cp = d.argmin(1)
    print("  best matching want-col for got cols 0..15:", cp[:16].tolist(), " min dist", float(d.min(1).values.max()))
    # error map per 32x32 block
    err = (gm - wm).abs()
    blocks = err.reshape(4, 32, 8, 32).amax(dim=(1, 3))
    print("  max abs err per 32x32 block (rows x cols):")
    for r in range(4):
        print("   ", [f"{v:.2e}" for v in blocks[r].tolist()])
    return False


if __name__ == "__main__":
    ok = True
    for shape in [(128, 256, 16), (128, 256, 64), (128, 256, 128), (256, 512, 320)]:
        ok = probe(*shape) and ok
    print("[probe] all ok" if ok else "[probe] FAILED")
