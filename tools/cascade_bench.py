"""All the GPU stages this repository provides, chained the way the reference cascade chains them
(models/sam_maskdecoder_edge.py infer_test :333-357, cocotrainers/mapleAlphaCLIP.py:281-296): raw uint8 images ->
preprocess (resize + normalise) -> SAM-Adapter ViT-H encoder;  Alpha-CLIP tower on the 336^2 crop with an all-ones alpha map
and again with a mask -> (the CLIP text tower and the two prompt projections are the reference's: synthetic prompt tokens
here) -> mask decoder -> postprocess.  Resident inputs, CUDA events, one JSON line.

    python tools/cascade_bench.py [--batch 32] [--steps 5] [--warmup 3]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from camouflaged_vlm_b200 import PRESETS, build_encoder                     # noqa: E402
from camouflaged_vlm_b200 import mask_decoder as md                         # noqa: E402
from camouflaged_vlm_b200.preprocess import preprocess_u8                    # noqa: E402
from camouflaged_vlm_b200.synth import seeded_state_dict                    # noqa: E402
from tower_bench import build_tower                                         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    B = args.batch
    cfg = PRESETS["vit_h"]
    enc = build_encoder(cfg).eval()
    enc.load_state_dict(seeded_state_dict(cfg, 0), strict=True)
    enc = enc.cuda()
    dec = md.build_mask_decoder(seed=0).cuda()
    tower = build_tower()
    gen = torch.Generator().manual_seed(0)
    raw = [torch.randint(0, 256, (768, 1024, 3), generator=gen, dtype=torch.uint8).cuda() for _ in range(B)]
    clip_img = torch.randn(B, 3, 336, 336, generator=gen).cuda()
    ones = torch.ones(B, 1, 336, 336).cuda()
    mask = (torch.rand(B, 1, 336, 336, generator=gen) > 0.5).float().cuda()
    shared = (torch.randn(2, 1024, generator=gen) * 0.02).cuda()
    deep = [(torch.randn(2, 1024, generator=gen) * 0.02).cuda() for _ in range(8)]
    pe = torch.randn(1, 256, 64, 64, generator=gen).cuda()
    dense = (torch.randn(1, 256, 1, 1, generator=gen) * 0.05).cuda().expand(B, -1, 64, 64)
    proj = (torch.randn(768, 256, generator=gen) * 0.03).cuda()          # stand-in for sam_visual_proj / sam_text_proj
    stages = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(record=False):
        marks = [ev() for _ in range(5)] if record else None
        with torch.no_grad():
            if record: marks[0].record()
            x = preprocess_u8(raw, size=1024)
            if record: marks[1].record()
            feats, interm = enc(x, interm=True)
            if record: marks[2].record()
            f1 = tower(clip_img, ones, shared, deep)
            f2 = tower(clip_img, mask, shared, deep)
            if record: marks[3].record()
            sparse = torch.stack([f1 @ proj, f2 @ proj], dim=1)          # [B, 2, 256]: image / text feature stand-ins
            m, e, i = dec(feats, interm, pe, sparse, dense, False)
            out = md.postprocess_masks(m, 1024, 1024, 1024)
            if record: marks[4].record()
        return marks, out

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    s, e = ev(), ev()
    s.record()
    for _ in range(args.steps):
        step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.steps
    marks, out = step(record=True)
    torch.cuda.synchronize()
    names = ("preprocess", "encoder", "clip_tower_x2", "decoder_postprocess")
    stages = {n: marks[i].elapsed_time(marks[i + 1]) for i, n in enumerate(names)}
    print(json.dumps({"metric": "cascade_gpu_stages_img_per_s", "value": B / ms * 1e3, "unit": "img/s", "ms_per_step": ms,
                      "batch": B, "steps": args.steps, "warmup": args.warmup, "stages_ms": stages,
                      "gpu_launches": enc.last_launch_count + 2 * tower.last_launch_count + dec.last_launch_count + 2 * B + 1,
                      "config": {"workload": "768x1024 uint8 images -> preprocess -> ViT-H encoder (interm) ; ViT-L/14@336 tower "
                                             "twice ; mask decoder -> 1024^2 logits; CLIP text tower / prompt projections not included",
                                 "out": list(out.shape)}}))


if __name__ == "__main__":
    main()
