"""Sustained (power-capped) throughput of single kernels: each is launched back to back for a few seconds and timed
over the last part, with nvidia-smi clocks / power sampled meanwhile -- at the cap a kernel's sustained rate IS its
energy efficiency.  The four block GEMMs of ViT-H next to torch.matmul (cuBLAS) on the same shapes, and the two
attention kernels.  GPU only.   python tools/sustained_probe.py [batch] [seconds]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import ClockSampler
from camouflaged_vlm_b200 import ops


def sustained(fn, seconds, flops):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds * 0.4:       # heat up: not timed
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds * 0.6:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / n
    return {"ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1), "sm_mhz": clocks.get("sm_mhz"),
            "power_w_max": clocks.get("power_w_max"), "reasons": clocks.get("reasons")}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
    M, D, H, hd = B * 4096, 1280, 16, 80
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device="cuda") * sc)
    res = rn(M, D)
    out = {"batch": B, "seconds": secs}
    for name, n, k, kw in (("qkv", 3 * D, D, dict(out_dtype=torch.bfloat16)),
                           ("proj", D, D, dict(residual=res)),
                           ("lin1", 4 * D, D, dict(out_dtype=torch.bfloat16, gelu=True)),
                           ("lin2", D, 4 * D, dict(residual=res))):
        x = rn(M, k).bfloat16()
        w = rn(n, k, sc=0.02).bfloat16()
        bias = rn(n, sc=0.02)
        o = torch.empty(M, n, device="cuda", dtype=kw.get("out_dtype", torch.float32))
        kw2 = {kk: vv for kk, vv in kw.items() if kk != "out_dtype"}
        fl = 2.0 * M * n * k
        ours = sustained(lambda: ops.gemm_bf16(x, w, bias=bias, out=o, **kw2), secs, fl)
        ob = torch.empty(M, n, device="cuda", dtype=torch.bfloat16)
        wt = w.t()
        cublas = sustained(lambda: torch.matmul(x, wt, out=ob), secs, fl)      # plain bf16 GEMM, no epilogue work
        out[name] = {"ours": ours, "torch_matmul_bf16": cublas}
        print(name, json.dumps(out[name]), flush=True)
        del x, w, o, ob
    qkv = rn(M, 3 * D, sc=0.7).bfloat16()
    for name, window, s in (("attn_global", 0, 64), ("attn_window", 14, 14)):
        relh = rn(2 * s - 1, hd, sc=0.1).bfloat16()
        relw = rn(2 * s - 1, hd, sc=0.1).bfloat16()
        keys = 4096 if window == 0 else 196
        fl = B * H * (4.0 * 4096 * keys * hd + 2.0 * 4096 * 2 * s * hd)
        out[name] = sustained(lambda: ops.attention(qkv, relh, relw, B, 64, window, H, hd), secs, fl)
        print(name, json.dumps(out[name]), flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
