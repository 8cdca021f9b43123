"""Alpha-CLIP / MaPLe visual tower timing on one GPU: ViT-L/14@336 with 2 context tokens and 8 deeper
prompts, images resident in HBM.  CUDA events on the launching stream after warm-up; the activations of a batch (> 1 GB)
exceed L2.  Prints one JSON line; the CPU baseline beside it (the oracle on the host cores) is measured by
tests/cpu_baselines.py: only tests/, smoke() and bench.py may run anything under oracle/.

    python tools/tower_bench.py [--batch 32] [--steps 10] [--warmup 3]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from camouflaged_vlm_b200 import alpha_clip as ac      # noqa: E402


def build_tower(seed=0, width=1024, layers=24, heads=16, res=336, out_dim=768, n_ctx=2):
    m = ac.VisionTransformer_MaPLe(res, 14, width, layers, heads, out_dim, design_details={"trainer": "MaPLe", "maple_length": n_ctx}).eval()
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in m.named_parameters():
            std = 0.5 if n.startswith("conv1") else width ** -0.5
            p.copy_(torch.randn(p.shape, generator=gen) * std + (1.0 if (p.dim() == 1 and n.endswith("weight")) else 0.0))
    return m.cuda()


def flops_per_image(L=579, w=1024, layers=24, grid=24, patch=14):
    return layers * (2.0 * L * w * 12 * w + 4.0 * L * L * w) + 2.0 * grid * grid * w * 4 * patch * patch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    B = args.batch
    m = build_tower()
    gen = torch.Generator().manual_seed(1)
    img = torch.randn(B, 3, 336, 336, generator=gen).cuda()
    alpha = (torch.rand(B, 1, 336, 336, generator=gen) > 0.5).float().cuda()
    shared = (torch.randn(2, 1024, generator=gen) * 0.02).cuda()
    deep = [(torch.randn(2, 1024, generator=gen) * 0.02).cuda() for _ in range(8)]

    def step():
        with torch.no_grad():
            return m(img, alpha, shared, deep)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.steps
    peak, src = 1391.4, "fallback: the sustained figure MEASURED_PEAKS.json held in round 1"
    try:
        pj = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        peak, src = float(pj["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json), sustained figure"
    except Exception:
        pass
    tf = B * flops_per_image() / (ms * 1e-3) / 1e12
    line = {"metric": "alpha_clip_vit_l14_336_tower_img_per_s", "value": B / ms * 1e3, "unit": "img/s", "ms_per_step": ms,
            "batch": B, "steps": args.steps, "warmup": args.warmup, "gpu_launches": m.last_launch_count, "dtype": "bf16",
            "config": {"workload": "VisionTransformer_MaPLe ViT-L/14@336, 579 tokens (2 context tokens, 8 deeper prompts), one "
                                   "pass (the cascade runs two per image)", "l2": "activations of a batch exceed L2"},
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                         "peak_source": src, "traffic": None, "flops_per_image": flops_per_image()}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
