"""Parity of the CUDA mask decoder (nn.Module -> C ABI -> csrc/decoder.cu) on the B200.

Checkers: the golden outputs of the reference's own MaskDecoder_Edge (tests/golden/decoder_small.npz, made by
oracle/make_golden.py) and the pinned oracle (oracle/mask_decoder_oracle.py) on seeded inputs.  The image-side GEMMs run
with bf16 operands, so the gates are the ones the encoder uses: relative L2 <= 1e-2 on every output and binary-mask
IoU >= 0.99; the token side is fp32."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from camouflaged_vlm_b200 import mask_decoder as md      # noqa: E402
from oracle import mask_decoder_oracle as mdo            # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-2


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def build(seed=0):
    sd = mdo.seeded_decoder_state(seed)
    dec = md.MaskDecoder(num_multimask_outputs=3, transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256,
                         transformer=md.TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8)).eval()
    dec.load_state_dict({k: v for k, v in sd.items() if not k.startswith("probe.")}, strict=True)
    return dec.cuda(), sd


def run(dec, sd, emb, sparse=None, multimask=False):
    B, _, g, _ = emb.shape
    sparse = sd["probe.sparse"].expand(B, -1, -1) if sparse is None else sparse
    dense = sd["probe.no_mask_embed"].reshape(1, -1, 1, 1).expand(B, -1, g, g)
    with torch.no_grad():
        return dec(image_embeddings=emb.cuda(), interm_embeddings=[], image_pe=mdo.dense_pe(sd, g).cuda(),
                   sparse_prompt_embeddings=sparse.cuda(), dense_prompt_embeddings=dense.cuda(), multimask_output=multimask)


def test_decoder_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "decoder_small.npz"))
    dec, sd = build(0)
    emb = torch.from_numpy(g["emb"])
    masks, edge, iou = run(dec, sd, emb)
    assert masks.shape == (1, 1, 64, 64) and edge.shape == (1, 1, 64, 64) and iou.shape == (1, 1)
    want = torch.from_numpy(g["logits"])
    assert rel_l2(masks.cpu(), want) < TOL
    assert rel_l2(edge.cpu(), torch.from_numpy(g["edge"])) < TOL
    assert float((iou.cpu() - torch.from_numpy(g["iou"])[:, 0:1]).abs().max()) < 1e-3
    assert mdo.iou(mdo.binary_mask(masks.cpu(), 256), mdo.binary_mask(want, 256)) >= 0.99
    m3, _, i3 = run(dec, sd, emb, multimask=True)                  # mask_decoder_edge.py:133-139: tokens 1..3
    assert m3.shape == (1, 3, 64, 64) and i3.shape == (1, 3)
    assert rel_l2(m3.cpu(), torch.from_numpy(g["masks_all"])[:, 1:]) < TOL
    assert dec.last_launch_count == 37


def test_decoder_token_side_is_fp32_accurate():
    # hs (the transformer's output tokens) goes through fp32 kernels except for the bf16 key / value projections
    dec, sd = build(1)
    emb = torch.randn(1, 256, 16, 16, generator=torch.Generator().manual_seed(5))
    run(dec, sd, emb)
    want = mdo.decode_all(sd, emb)["hs"]
    assert rel_l2(dec.last_hs.cpu(), want) < 3e-3


def test_decoder_batched_equals_per_image_oracle():
    dec, sd = build(2)
    gen = torch.Generator().manual_seed(11)
    emb = torch.randn(3, 256, 16, 16, generator=gen)
    sparse = torch.randn(3, 2, 256, generator=gen)
    masks, edge, iou = run(dec, sd, emb, sparse)
    for b in range(3):
        want = mdo.decode_all(sd, emb[b:b + 1], sparse[b:b + 1])
        assert rel_l2(masks[b:b + 1].cpu(), want["masks"][:, 0:1]) < TOL, b
        assert rel_l2(edge[b:b + 1].cpu(), want["edge"]) < TOL, b
        assert float((iou[b:b + 1].cpu() - want["iou"][:, 0:1]).abs().max()) < 1e-3
        assert mdo.iou(mdo.binary_mask(masks[b:b + 1].cpu(), 256), mdo.binary_mask(want["masks"][:, 0:1], 256)) >= 0.99
        # the result for an image does not depend on what else is in the batch, nor on the chunking
        one = run(dec, sd, emb[b:b + 1], sparse[b:b + 1])
        assert torch.equal(one[0], masks[b:b + 1]) and torch.equal(one[1], edge[b:b + 1])
    dec.max_chunk = 2
    again = run(dec, sd, emb, sparse)
    assert torch.equal(again[0], masks) and torch.equal(again[2], iou)


def test_decoder_full_size_and_dense_per_image():
    # the deployed geometry: 64 x 64 embedding -> 256 x 256 logits, per-image dense prompts, 3 prompt tokens
    dec, sd = build(0)
    gen = torch.Generator().manual_seed(2)
    emb = torch.randn(2, 256, 64, 64, generator=gen)
    sparse = torch.randn(2, 3, 256, generator=gen)
    dense = sd["probe.no_mask_embed"].reshape(1, -1, 1, 1).expand(2, -1, 64, 64).contiguous()
    with torch.no_grad():
        masks, edge, iou = dec(emb.cuda(), [], mdo.dense_pe(sd, 64).cuda(), sparse.cuda(), dense.cuda(), False)
    assert masks.shape == (2, 1, 256, 256)
    for b in range(2):
        want = mdo.decode_all(sd, emb[b:b + 1], sparse[b:b + 1])
        assert rel_l2(masks[b:b + 1].cpu(), want["masks"][:, 0:1]) < TOL
        assert rel_l2(edge[b:b + 1].cpu(), want["edge"]) < TOL
        assert mdo.iou(mdo.binary_mask(masks[b:b + 1].cpu(), 1024), mdo.binary_mask(want["masks"][:, 0:1], 1024)) >= 0.99


def test_decoder_matches_reference_vit_h_logits():
    # the reference cascade's low-res logits for the reference encoder's ViT-H embedding are committed (fp16); feed the
    # same embedding (recomputed by OUR encoder, itself within 1e-2) -- end-to-end encoder + decoder against the reference
    from camouflaged_vlm_b200 import PRESETS, build_encoder
    from camouflaged_vlm_b200.synth import seeded_images, seeded_state_dict
    cfg = PRESETS["vit_h"]
    enc = build_encoder(cfg).eval()
    enc.load_state_dict(seeded_state_dict(cfg, 0), strict=True)
    enc = enc.cuda()
    with torch.no_grad():
        emb = enc(seeded_images(cfg, 1).cuda())
    dec, sd = build(0)
    masks, _, _ = run(dec, sd, emb)
    want = torch.from_numpy(np.load(os.path.join(GOLDEN, "decoder_vit_h_logits.npz"))["logits"].astype(np.float32))
    assert mdo.iou(mdo.binary_mask(masks.cpu(), 1024), mdo.binary_mask(want, 1024)) >= 0.99


def test_bilinear_resize_and_postprocess():
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 64, 48, generator=gen).cuda()
    for size in ((256, 192), (100, 77), (64, 48), (31, 200)):
        got = md.bilinear_resize(x, size)
        want = F.interpolate(x, size, mode="bilinear", align_corners=False)
        assert float((got - want).abs().max()) < 2e-6, size
    m = torch.randn(1, 1, 256, 256, generator=gen).cuda()
    got = md.postprocess_masks(m, 1024, 1024, 1024)
    want = F.interpolate(F.interpolate(m, (1024, 1024), mode="bilinear", align_corners=False)[..., :1024, :1024], (1024, 1024),
                         mode="bilinear", align_corners=False)
    assert float((got - want).abs().max()) < 2e-6
    got = md.postprocess_masks(m, 1024, 900, (480, 640))
    want = F.interpolate(F.interpolate(m, (1024, 1024), mode="bilinear", align_corners=False)[..., :900, :900], (480, 640),
                         mode="bilinear", align_corners=False)
    assert float((got - want).abs().max()) < 2e-6
    assert ((got > 0) == (want > 0)).float().mean() > 0.9999


def test_decoder_rejects_bad_shapes():
    dec, sd = build(0)
    with torch.no_grad():
        with pytest.raises(ValueError):
            dec(torch.zeros(2, 256, 16, 16).cuda(), [], mdo.dense_pe(sd, 16).cuda(), torch.zeros(1, 2, 256).cuda(),
                torch.zeros(1, 256, 16, 16).cuda(), False)
        with pytest.raises(Exception, match="multiple of 8"):
            dec(torch.zeros(1, 256, 12, 12).cuda(), [], mdo.dense_pe(sd, 12).cuda(), torch.zeros(1, 2, 256).cuda(),
                torch.zeros(1, 256, 12, 12).cuda(), False)


@pytest.mark.parametrize("csize", [1, 2, 4, 8])
def test_decoder_token_cluster_sizes(csize, request):
    # the token side runs one CTA cluster per image; every cluster size must give the same answer (up to the fp32
    # summation order of the K split) and match the oracle
    from camouflaged_vlm_b200 import _lib
    _lib.set_option("dec_cluster", csize)
    request.addfinalizer(lambda: _lib.set_option("dec_cluster", 0))
    dec, sd = build(1)
    gen = torch.Generator().manual_seed(21)
    emb = torch.randn(2, 256, 16, 16, generator=gen)
    sparse = torch.randn(2, 3, 256, generator=gen)
    masks, edge, iou = run(dec, sd, emb, sparse)
    hs = dec.last_hs.clone()
    for b in range(2):
        want = mdo.decode_all(sd, emb[b:b + 1], sparse[b:b + 1])
        assert rel_l2(hs[b:b + 1].cpu(), want["hs"]) < 3e-3
        assert rel_l2(masks[b:b + 1].cpu(), want["masks"][:, 0:1]) < TOL
        assert float((iou[b:b + 1].cpu() - want["iou"][:, 0:1]).abs().max()) < 1e-3
    again = run(dec, sd, emb, sparse)
    assert torch.equal(again[0], masks) and torch.equal(dec.last_hs, hs)      # deterministic run to run


@pytest.mark.parametrize("g,nc,batch", [(8, 1, 1), (8, 8, 3), (32, 5, 2)])
def test_decoder_edge_geometries(g, nc, batch):
    # smallest embedding (64 keys: fewer keys per cluster rank than warps), one and the maximum number of prompt tokens
    dec, sd = build(4)
    gen = torch.Generator().manual_seed(100 + g + nc)
    emb = torch.randn(batch, 256, g, g, generator=gen)
    sparse = torch.randn(batch, nc, 256, generator=gen)
    masks, edge, iou = run(dec, sd, emb, sparse, multimask=True)
    assert masks.shape == (batch, 3, 4 * g, 4 * g) and edge.shape == (batch, 1, 4 * g, 4 * g) and iou.shape == (batch, 3)
    for b in range(batch):
        want = mdo.decode_all(sd, emb[b:b + 1], sparse[b:b + 1])
        assert rel_l2(masks[b:b + 1].cpu(), want["masks"][:, 1:]) < TOL
        assert rel_l2(edge[b:b + 1].cpu(), want["edge"]) < TOL
        assert rel_l2(dec.last_hs[b:b + 1].cpu(), want["hs"]) < 3e-3
        assert float((iou[b:b + 1].cpu() - want["iou"][:, 1:]).abs().max()) < 1e-3


def test_decoder_repacks_after_load_state_dict():
    dec, sd = build(0)
    emb = torch.randn(1, 256, 16, 16, generator=torch.Generator().manual_seed(9))
    a = run(dec, sd, emb)[0].clone()
    sd9 = mdo.seeded_decoder_state(9)
    dec.load_state_dict({k: v for k, v in sd9.items() if not k.startswith("probe.")}, strict=True)     # must invalidate the packed copies
    b = run(dec, sd, emb)[0].clone()
    dec.load_state_dict({k: v for k, v in sd.items() if not k.startswith("probe.")}, strict=True)
    c = run(dec, sd, emb)[0]
    assert rel_l2(b, a) > 1e-2 and torch.equal(a, c)
    import copy
    twin = copy.deepcopy(dec)                                                                           # owns its own native handle
    assert torch.equal(run(twin, sd, emb)[0], a)


def test_graphed_decoder_and_tower_replay_bit_identical():
    from camouflaged_vlm_b200.graphed import GraphedCall
    from camouflaged_vlm_b200 import alpha_clip as ac
    from oracle import alpha_clip_oracle as aco
    dec, sd = build(0)
    gen = torch.Generator().manual_seed(31)
    emb = [torch.randn(1, 256, 16, 16, generator=gen).cuda() for _ in range(2)]
    sparse = [torch.randn(1, 2, 256, generator=gen).cuda() for _ in range(2)]
    pe = mdo.dense_pe(sd, 16).cuda()
    dense = sd["probe.no_mask_embed"].reshape(1, -1, 1, 1).expand(1, -1, 16, 16).contiguous().cuda()
    fn = lambda e, s: dec(e, [], pe, s, dense, False)
    g = GraphedCall(fn, emb[0], sparse[0])
    for e, s in zip(emb, sparse):
        with torch.no_grad():
            want = fn(e, s)
        got = g(e, s)
        assert all(torch.equal(a, b) for a, b in zip(got, want))
    cfg = aco.SMALL_HD64
    tw = ac.VisionTransformer_MaPLe(cfg.input_resolution, cfg.patch_size, cfg.width, cfg.layers, cfg.heads, cfg.output_dim,
                                    design_details={"maple_length": cfg.n_ctx}).eval()
    tw.load_state_dict(aco.seeded_vit_state(cfg, 0), strict=True)
    tw = tw.cuda()
    shared, deep = aco.seeded_prompts(cfg, depth=3)
    shared, deep = shared.cuda(), [d.cuda() for d in deep]
    x = [torch.randn(1, 3, cfg.input_resolution, cfg.input_resolution, generator=gen).cuda() for _ in range(2)]
    al = torch.ones(1, 1, cfg.input_resolution, cfg.input_resolution).cuda()
    tfn = lambda xi, ai: tw(xi, ai, shared, deep)
    tg = GraphedCall(tfn, x[0], al)
    for xi in x:
        with torch.no_grad():
            want = tfn(xi, al)
        assert torch.equal(tg(xi, al), want)


def test_graph_survives_workspace_growth_and_detects_weight_change():
    """a captured graph has the module's workspace and packed weights baked in.  A later eager call with a
    larger batch replaces the module's workspace -- the graph must keep working on its own (kept-alive) buffer -- and a
    parameter update must make the replay raise instead of running on stale packed weights."""
    from camouflaged_vlm_b200.graphed import GraphedCall
    dec, sd = build(0)
    gen = torch.Generator().manual_seed(41)
    emb = torch.randn(1, 256, 16, 16, generator=gen).cuda()
    sparse = torch.randn(1, 2, 256, generator=gen).cuda()
    pe = mdo.dense_pe(sd, 16).cuda()
    dense = sd["probe.no_mask_embed"].reshape(1, -1, 1, 1).expand(1, -1, 16, 16).contiguous().cuda()
    fn = lambda e, s: dec(e, [], pe, s, dense, False)
    g = GraphedCall(fn, emb, sparse)
    assert g._owners and g._owners[0] is dec                       # found through the lambda's closure
    with torch.no_grad():
        want = [t.clone() for t in fn(emb, sparse)]
        ws_before = dec._workspace.data_ptr()
        big = torch.randn(6, 256, 16, 16, generator=gen).cuda()
        dec(big, [], pe, torch.randn(6, 2, 256, generator=gen).cuda(), dense, False)      # needs a larger workspace
        assert dec._workspace.data_ptr() != ws_before
        junk = torch.full((dec._workspace.numel() // 4,), float("nan"), device="cuda")     # churn the allocator
        del junk
    got = g(emb, sparse)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(got, want))
    with torch.no_grad():
        next(dec.parameters()).add_(1.0)
    with pytest.raises(RuntimeError, match="parameters changed"):
        g(emb, sparse)
