#!/usr/bin/env python
"""Workload for tools/gpu_sanitize.sh (compute-sanitizer memcheck / racecheck / synccheck): one pass over every kernel
family at small shapes -- the encoder on the small fixtures (mma.sync attention, 14x14 tcgen05 window attention, CTA-pair
and single-CTA GEMMs, norm1 fold, native high-pass), a narrow 1024^2 encoder so that the 64x64 tcgen05 global-attention
kernel runs, the mask decoder (cluster token kernels, implicit-GEMM convolutions), the CLIP tower (plain tcgen05
attention) and the input pipeline.  Results are checked loosely (finite + oracle rel-L2) so a silent corruption also fails."""
import os
import sys
from dataclasses import replace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from camouflaged_vlm_b200 import PRESETS, build_encoder
from camouflaged_vlm_b200.synth import seeded_images, seeded_state_dict
from oracle import encoder_oracle as orc


def encoder_case(cfg, batch, name, check=True):
    sd = seeded_state_dict(cfg, 0)
    m = build_encoder(cfg).eval()
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    x = seeded_images(cfg, batch, 0)
    with torch.no_grad():
        out, interm = m(x.cuda(), interm=True)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    if check:
        err = orc.rel_l2(out.cpu(), orc.encoder_forward(sd, cfg, x))
        assert err < 1e-2, (name, err)
        print(f"[sanitize] {name}: rel-L2 {err:.2e}, {m.last_launch_count} launches", flush=True)
    else:
        print(f"[sanitize] {name}: finite, {m.last_launch_count} launches", flush=True)


def main():
    which = sys.argv[1:] or ["tiny", "small", "global", "decoder", "tower", "preprocess", "train"]
    if "tiny" in which:
        encoder_case(PRESETS["tiny"], 2, "encoder tiny B=2")
    if "small" in which:
        encoder_case(PRESETS["small_hd80"], 1, "encoder small_hd80 B=1 (14x14 tcgen05 windows)")
        encoder_case(PRESETS["small_hd64"], 1, "encoder small_hd64 B=1")
    if "global" in which:
        # 64x64 token grid, 2 heads x 80: the tcgen05 global-attention kernel + full-size window kernel geometry
        cfg = replace(PRESETS["vit_h"], embed_dim=160, depth=2, num_heads=2, global_attn_indexes=(1,))
        encoder_case(cfg, 1, "encoder 1024^2 D=160 depth 2 (64x64 tcgen05 global attention)", check=True)
    if "decoder" in which:
        from camouflaged_vlm_b200 import mask_decoder as md
        from oracle import mask_decoder_oracle as mdo
        dsd = mdo.seeded_decoder_state(0)
        dec = md.build_mask_decoder()
        dec.load_state_dict({k: v for k, v in dsd.items() if not k.startswith("probe.")}, strict=True)
        dec = dec.cuda()
        emb = torch.randn(2, 256, 16, 16, generator=torch.Generator().manual_seed(0))
        sparse = torch.randn(2, 2, 256, generator=torch.Generator().manual_seed(1))
        dense = dsd["probe.no_mask_embed"].reshape(1, -1, 1, 1).expand(2, -1, 16, 16)
        with torch.no_grad():
            masks, edge, iou = dec(emb.cuda(), [], mdo.dense_pe(dsd, 16).cuda(), sparse.cuda(), dense.cuda(), False)
            up = md.postprocess_masks(masks, 256, 256, 256)
        torch.cuda.synchronize()
        ref = mdo.decode_all(dsd, emb[0:1], sparse[0:1])
        err = orc.rel_l2(masks[0:1].cpu(), ref["masks"][:, 0:1])
        assert err < 1e-2 and torch.isfinite(up).all(), err
        print(f"[sanitize] mask decoder B=2: rel-L2 {err:.2e}, {dec.last_launch_count} launches", flush=True)
    if "tower" in which:
        from camouflaged_vlm_b200 import alpha_clip as ac
        from oracle import alpha_clip_oracle as aco
        vcfg = aco.SMALL_HD64
        vsd = aco.seeded_vit_state(vcfg, 0)
        tower = ac.VisionTransformer_MaPLe(vcfg.input_resolution, vcfg.patch_size, vcfg.width, vcfg.layers, vcfg.heads,
                                           vcfg.output_dim, design_details={"trainer": "MaPLe", "maple_length": vcfg.n_ctx}).eval()
        tower.load_state_dict(vsd, strict=True)
        tower = tower.cuda()
        shared, deep = aco.seeded_prompts(vcfg, depth=3)
        gen = torch.Generator().manual_seed(0)
        vimg = torch.randn(2, 3, vcfg.input_resolution, vcfg.input_resolution, generator=gen)
        valpha = (torch.rand(2, 1, vcfg.input_resolution, vcfg.input_resolution, generator=gen) > 0.5).float()
        with torch.no_grad():
            feat = tower(vimg.cuda(), valpha.cuda(), shared.cuda(), [d.cuda() for d in deep])
        torch.cuda.synchronize()
        err = orc.rel_l2(feat.cpu(), aco.vit_maple_forward(vsd, vcfg, vimg, valpha, shared, deep))
        assert err < 1e-2, err
        print(f"[sanitize] CLIP tower B=2: rel-L2 {err:.2e}, {tower.last_launch_count} launches", flush=True)
    if "preprocess" in which:
        from camouflaged_vlm_b200.preprocess import preprocess_u8
        img = torch.randint(0, 256, (300, 420, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3)).cuda()
        out = preprocess_u8(img, size=256)
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()
        print("[sanitize] preprocess 300x420 -> 256^2 ok", flush=True)
    if "train" in which:
        try:
            from camouflaged_vlm_b200 import backward  # noqa: F401
        except ImportError:
            print("[sanitize] train step: backward not built, skipped", flush=True)
        else:
            cfg = PRESETS["small_hd80"]
            m = build_encoder(cfg)
            m.load_state_dict(seeded_state_dict(cfg, 0), strict=True)
            m = m.cuda().train()
            for n, p in m.named_parameters():
                p.requires_grad_("prompt_generator" in n)
            out = m(seeded_images(cfg, 1, 0).cuda())
            out.square().mean().backward()
            torch.cuda.synchronize()
            g = [p.grad for n, p in m.named_parameters() if p.requires_grad]
            assert all(t is not None and torch.isfinite(t).all() for t in g)
            print(f"[sanitize] adapter train step small_hd80: {len(g)} gradients finite", flush=True)
    print("[sanitize] done", flush=True)


if __name__ == "__main__":
    main()
