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


def test_b200_module_parameter_tree_matches_reference_shapes():
    from camouflaged_vlm_b200 import alpha_clip as ac
    cfg = aco.SMALL_HD64
    m = ac.VisionTransformer_MaPLe(cfg.input_resolution, cfg.patch_size, cfg.width, cfg.layers, cfg.heads, cfg.output_dim,
                                   design_details={"trainer": "MaPLe", "maple_length": cfg.n_ctx})
    shapes = aco.vit_param_shapes(cfg)
    sd = m.state_dict()
    assert set(sd) == set(shapes)
    for k, v in sd.items():
        assert tuple(v.shape) == shapes[k], k
    m.load_state_dict(aco.seeded_vit_state(cfg, 0), strict=True)
    packed = ac.pack_vit_weights(dict(m.state_dict()), cfg.layers, torch.device("cpu"))
    assert packed["patch.w"].shape == (cfg.width, 4 * 14 * 14) and packed["proj.t"].shape == (cfg.output_dim, cfg.width)
    # the patch GEMM's weight is conv1 and conv1_alpha side by side
    x = torch.randn(1, 3, cfg.input_resolution, cfg.input_resolution)
    a = torch.randn(1, 1, cfg.input_resolution, cfg.input_resolution)
    want = torch.nn.functional.conv2d(x, m.conv1.weight, stride=14) + torch.nn.functional.conv2d(a, m.conv1_alpha.weight, stride=14)
    cols = torch.nn.functional.unfold(torch.cat([x, a], 1), 14, stride=14)[0].t()            # [G*G, 4*196], (c, py, px)
    got = (cols @ torch.cat([m.conv1.weight.reshape(cfg.width, -1), m.conv1_alpha.weight.reshape(cfg.width, -1)], 1).t())
    assert float((got.t().reshape(want.shape) - want).abs().max()) < 1e-4
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):
        m(x, a, torch.zeros(cfg.n_ctx, cfg.width), [])
