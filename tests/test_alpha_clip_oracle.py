"""Pin the Alpha-CLIP / MaPLe visual tower oracle (groundwork for the next row: no CUDA path yet) against
the reference module: committed golden outputs, plus the live module when /root/reference is present."""
import os

import numpy as np
import pytest
import torch

from oracle import alpha_clip_oracle as aco
from oracle.ref_import import reference_available

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "alpha_clip_small.npz"))
    cfg = aco.SMALL
    sd = aco.seeded_vit_state(cfg, 0)
    shared, deep = aco.seeded_prompts(cfg, depth=3)
    image, alpha = torch.from_numpy(g["image"]), torch.from_numpy(g["alpha"])
    got = aco.vit_maple_forward(sd, cfg, image, alpha, shared, deep)
    assert got.shape == (2, cfg.output_dim)
    assert float((got - torch.from_numpy(g["out"])).abs().max()) < 1e-5
    got_shallow = aco.vit_maple_forward(sd, cfg, image, alpha, shared, [])
    assert float((got_shallow - torch.from_numpy(g["out_shallow"])).abs().max()) < 1e-5
    assert float((got - got_shallow).abs().max()) > 1e-4          # the deep prompts do change the features


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference_keys_and_output():
    from oracle.make_golden_alpha_clip import reference_tower, seeded_inputs
    cfg = aco.VitConfig(input_resolution=42, patch_size=14, width=48, layers=4, heads=3, output_dim=16, n_ctx=3)
    sd = aco.seeded_vit_state(cfg, 5)
    vit = reference_tower(cfg, sd)
    ref = vit.state_dict()
    shapes = aco.vit_param_shapes(cfg)
    assert set(ref) == set(shapes)
    for k, v in ref.items():
        assert tuple(v.shape) == shapes[k], k
    image, alpha = seeded_inputs(cfg, 3, seed=5)
    shared, deep = aco.seeded_prompts(cfg, depth=9, seed=5)        # more deep prompts than layers: the extra ones are unused
    with torch.no_grad():
        want = vit(image, alpha, shared, deep)
    assert float((aco.vit_maple_forward(sd, cfg, image, alpha, shared, deep) - want).abs().max()) < 1e-5


def test_full_size_geometry():
    cfg = aco.VitConfig()
    assert cfg.grid == 24 and cfg.tokens == 579
    assert 3.5e11 < aco.flops_per_image(cfg) < 4.2e11            # ~0.38 TFLOP per pass, two passes per image in the cascade
