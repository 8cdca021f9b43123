"""Golden vectors for the Alpha-CLIP / MaPLe visual tower oracle: the UNMODIFIED reference module
(alpha_clip_rw/model.py VisionTransformer_MaPLe) on seeded weights and inputs.  Container only (needs /root/reference).

    python -m oracle.make_golden_alpha_clip
"""
import os

import numpy as np
import torch

from oracle import alpha_clip_oracle as aco
from oracle.ref_import import load_reference_alpha_clip_model

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def reference_tower(cfg, sd):
    m = load_reference_alpha_clip_model()
    vit = m.VisionTransformer_MaPLe(input_resolution=cfg.input_resolution, patch_size=cfg.patch_size, width=cfg.width,
                                    layers=cfg.layers, heads=cfg.heads, output_dim=cfg.output_dim,
                                    design_details={"trainer": "MaPLe", "maple_length": cfg.n_ctx}).eval()
    vit.load_state_dict(sd, strict=True)
    return vit


def seeded_inputs(cfg, batch, seed=0):
    gen = torch.Generator().manual_seed(seed)
    image = torch.randn(batch, 3, cfg.input_resolution, cfg.input_resolution, generator=gen)
    alpha = (torch.rand(batch, 1, cfg.input_resolution, cfg.input_resolution, generator=gen) > 0.5).float()
    return image, alpha


def main():
    for name, cfg, depth in (("small", aco.SMALL, 3), ("small_hd64", aco.SMALL_HD64, 3)):
        sd = aco.seeded_vit_state(cfg, 0)
        vit = reference_tower(cfg, sd)
        image, alpha = seeded_inputs(cfg, 2)
        shared, deep = aco.seeded_prompts(cfg, depth=depth)
        with torch.no_grad():
            out = vit(image, alpha, shared, deep)
            out_shallow = vit(image, alpha, shared, [])
        np.savez_compressed(os.path.join(GOLDEN, f"alpha_clip_{name}.npz"), image=image.numpy(), alpha=alpha.numpy(),
                            out=out.numpy(), out_shallow=out_shallow.numpy())
        print(f"[golden] alpha_clip_{name}", tuple(out.shape), float(out.std()))


if __name__ == "__main__":
    main()
