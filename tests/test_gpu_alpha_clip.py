"""Parity of the CUDA Alpha-CLIP / MaPLe visual tower (nn.Module -> C ABI -> csrc/alpha_clip.cu) on the B200
(SURVEY.md 8 f2).  Checkers: golden outputs of the reference's own VisionTransformer_MaPLe
(tests/golden/alpha_clip_small_hd64.npz, oracle/make_golden_alpha_clip.py) and the pinned oracle on seeded inputs at the
deployed ViT-L/14@336 geometry.  bf16 tensor-core operands, fp32 residual stream: relative L2 <= 1e-2 like the encoder."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from camouflaged_vlm_b200 import alpha_clip as ac        # noqa: E402
from oracle import alpha_clip_oracle as aco              # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-2


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def build(cfg, seed=0):
    sd = aco.seeded_vit_state(cfg, seed)
    m = ac.VisionTransformer_MaPLe(cfg.input_resolution, cfg.patch_size, cfg.width, cfg.layers, cfg.heads, cfg.output_dim,
                                   design_details={"trainer": "MaPLe", "maple_length": cfg.n_ctx}).eval()
    m.load_state_dict(sd, strict=True)
    return m.cuda(), sd


def test_tower_matches_reference_golden():
    cfg = aco.SMALL_HD64
    g = np.load(os.path.join(GOLDEN, "alpha_clip_small_hd64.npz"))
    m, sd = build(cfg)
    shared, deep = aco.seeded_prompts(cfg, depth=3)
    image, alpha = torch.from_numpy(g["image"]).cuda(), torch.from_numpy(g["alpha"]).cuda()
    with torch.no_grad():
        out = m(image, alpha, shared.cuda(), [d.cuda() for d in deep])
        out_shallow = m(image, alpha, shared.cuda(), [])
    assert out.shape == (2, cfg.output_dim) and out.dtype == torch.float32
    assert rel_l2(out.cpu(), torch.from_numpy(g["out"])) < TOL
    assert rel_l2(out_shallow.cpu(), torch.from_numpy(g["out_shallow"])) < TOL
    assert m.last_launch_count == 4 + 7 * cfg.layers + 3          # the shallow run: no context-token replacement launches


def test_tower_vit_l_336_matches_oracle_and_is_batch_invariant():
    cfg = aco.VitConfig()                                     # ViT-L/14@336, 577 + 2 tokens
    m, sd = build(cfg, seed=1)
    shared, deep = aco.seeded_prompts(cfg, depth=9, seed=1)   # MaPLe PROMPT_DEPTH 9: 8 deeper prompts
    gen = torch.Generator().manual_seed(3)
    image = torch.randn(3, 3, 336, 336, generator=gen)
    alpha = (torch.rand(3, 1, 336, 336, generator=gen) > 0.5).float()
    with torch.no_grad():
        out = m(image.cuda(), alpha.cuda(), shared.cuda(), [d.cuda() for d in deep])
        one = m(image[1:2].cuda(), alpha[1:2].cuda(), shared.cuda(), [d.cuda() for d in deep])
    want = aco.vit_maple_forward(sd, cfg, image[:2], alpha[:2], shared, deep)
    assert rel_l2(out[:2].cpu(), want) < TOL
    cos = torch.nn.functional.cosine_similarity(out[:2].cpu(), want, dim=-1)
    assert float(cos.min()) > 0.9999                          # what the cascade consumes: normalised features
    assert torch.equal(one, out[1:2])                         # independent of the rest of the batch
    m.max_chunk = 2
    with torch.no_grad():
        again = m(image.cuda(), alpha.cuda(), shared.cuda(), [d.cuda() for d in deep])
    assert torch.equal(again, out)


def test_tower_more_prompts_than_layers_and_half_inputs():
    cfg = aco.SMALL_HD64
    m, sd = build(cfg, seed=2)
    shared, deep = aco.seeded_prompts(cfg, depth=8, seed=2)   # 7 deeper prompts for 4 layers: the extra ones are unused
    gen = torch.Generator().manual_seed(5)
    image = torch.randn(5, 3, cfg.input_resolution, cfg.input_resolution, generator=gen)
    alpha = torch.ones(5, 1, cfg.input_resolution, cfg.input_resolution)
    want = aco.vit_maple_forward(sd, cfg, image, alpha, shared, deep)
    with torch.no_grad():
        out = m(image.cuda(), alpha.cuda(), shared.cuda(), [d.cuda() for d in deep])
        out_h = m(image.cuda().half(), alpha.cuda().half(), shared.cuda(), [d.cuda() for d in deep])
    assert rel_l2(out.cpu(), want) < TOL
    assert out_h.dtype == torch.float16 and rel_l2(out_h.float().cpu(), want) < 2e-2


def test_tower_rejects_bad_inputs():
    cfg = aco.SMALL_HD64
    m, _ = build(cfg)
    shared, _ = aco.seeded_prompts(cfg, depth=1)
    x = torch.zeros(1, 3, cfg.input_resolution, cfg.input_resolution).cuda()
    a = torch.zeros(1, 1, cfg.input_resolution, cfg.input_resolution).cuda()
    with torch.no_grad():
        with pytest.raises(NotImplementedError):
            m(x, a, shared.cuda(), [], return_attn=True)
        with pytest.raises(ValueError):
            m(x, a[:, :, :10], shared.cuda(), [])
        with pytest.raises(RuntimeError, match="no CPU path"):
            m(x.cpu(), a.cpu(), shared, [])
    with pytest.raises(Exception, match="head_dim"):
        bad = ac.VisionTransformer_MaPLe(56, 14, 128, 2, 8, 32, design_details={"maple_length": 2}).cuda()
        with torch.no_grad():
            bad(torch.zeros(1, 3, 56, 56).cuda(), torch.zeros(1, 1, 56, 56).cuda(), torch.zeros(2, 128).cuda(), [])


def test_tower_repacks_after_load_state_dict():
    cfg = aco.SMALL_HD64
    m, sd = build(cfg, seed=0)
    shared, deep = aco.seeded_prompts(cfg, depth=2)
    gen = torch.Generator().manual_seed(8)
    x = torch.randn(1, 3, cfg.input_resolution, cfg.input_resolution, generator=gen).cuda()
    al = torch.ones(1, 1, cfg.input_resolution, cfg.input_resolution).cuda()
    call = lambda mod: mod(x, al, shared.cuda(), [d.cuda() for d in deep])
    with torch.no_grad():
        a = call(m).clone()
        m.load_state_dict(aco.seeded_vit_state(cfg, 7), strict=True)        # must invalidate the packed copies
        b = call(m).clone()
        m.load_state_dict(sd, strict=True)
        c = call(m)
        import copy
        twin = copy.deepcopy(m)
        d = call(twin)
    assert rel_l2(b, a) > 1e-2 and torch.equal(a, c) and torch.equal(a, d)
