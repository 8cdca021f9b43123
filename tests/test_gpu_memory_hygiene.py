"""Memory hygiene of the C-ABI forwards without compute-sanitizer (closed on this GPU pool: profiles/sanitizer_closed_r2.log).

Every forward works inside one caller-provided workspace.  Two properties are checked for the encoder, the mask decoder
and the CLIP tower on the small fixtures (which run the tcgen05 window attention, CTA-pair and single-CTA GEMMs, the
cluster token kernels and the implicit-GEMM convolutions):

* **no read of uninitialised memory**: the outputs are bit-identical whether the workspace was pre-filled with zeros,
  with 0xFF bytes (bf16 / fp32 NaNs) or with large finite garbage -- a kernel that consumes a byte it did not write
  first (stale rows beyond a ragged tile, a missed zero-fill, a barrier that lets a consumer run early) changes bits;
* **no write outside the workspace**: a guard band of canary bytes after the bytes `*_workspace_bytes` asked for (and
  around the output tensors) is intact after the call.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from camouflaged_vlm_b200 import PRESETS, _lib, build_encoder           # noqa: E402
from camouflaged_vlm_b200.synth import seeded_images, seeded_state_dict  # noqa: E402

GUARD = 1 << 20
PATTERNS = (0x00, 0xFF, 0x7B)          # zeros, NaNs, large finite bf16 (0x7B7B ~ 1.3e36) / fp32 garbage


def _workspace_with_guard(need, fill):
    buf = torch.empty(need + GUARD, dtype=torch.uint8, device="cuda")
    buf[:need].fill_(fill)
    buf[need:].fill_(0xA5)
    return buf


def _guard_intact(buf, need):
    return bool((buf[need:] == 0xA5).all())


@pytest.mark.parametrize("name,batch", [("tiny", 2), ("small_hd80", 2), ("small_hd64", 1)])
def test_encoder_reads_nothing_it_did_not_write(name, batch):
    cfg = PRESETS[name]
    m = build_encoder(cfg).eval()
    m.load_state_dict(seeded_state_dict(cfg, 0), strict=True)
    m = m.cuda()
    x = seeded_images(cfg, batch, 0).cuda()
    with torch.no_grad():
        m(x)                                                   # creates the handle / packs the weights
    need = _lib.check(_lib.load().cvlm_encoder_workspace_bytes(m._handle, batch))
    results = []
    for fill in PATTERNS:
        ws = _workspace_with_guard(need, fill)
        m._workspace = ws[:need]                               # the module uses what it finds when it is large enough
        with torch.no_grad():
            out, interm = m(x, interm=True)
        torch.cuda.synchronize()
        assert m._workspace.data_ptr() == ws.data_ptr()
        assert _guard_intact(ws, need), f"{name}: write beyond the workspace (fill {fill:#x})"
        assert torch.isfinite(out).all()
        results.append((out.clone(), [t.clone() for t in interm]))
    for out, interm in results[1:]:
        assert torch.equal(out, results[0][0]), f"{name}: output depends on stale workspace contents"
        assert all(torch.equal(a, b) for a, b in zip(interm, results[0][1]))


def test_decoder_and_tower_read_nothing_they_did_not_write():
    from camouflaged_vlm_b200 import alpha_clip as ac
    from camouflaged_vlm_b200 import mask_decoder as md
    from oracle import alpha_clip_oracle as aco
    from oracle import mask_decoder_oracle as mdo
    dsd = mdo.seeded_decoder_state(0)
    dec = md.build_mask_decoder()
    dec.load_state_dict({k: v for k, v in dsd.items() if not k.startswith("probe.")}, strict=True)
    dec = dec.cuda()
    gen = torch.Generator().manual_seed(5)
    emb = torch.randn(3, 256, 16, 16, generator=gen).cuda()
    sparse = torch.randn(3, 2, 256, generator=gen).cuda()
    pe = mdo.dense_pe(dsd, 16).cuda()
    dense = dsd["probe.no_mask_embed"].reshape(1, -1, 1, 1).expand(3, -1, 16, 16).cuda()
    with torch.no_grad():
        dec(emb, [], pe, sparse, dense, False)
    need = dec._workspace.numel()
    outs = []
    for fill in PATTERNS:
        ws = _workspace_with_guard(need, fill)
        dec._workspace = ws[:need]
        with torch.no_grad():
            outs.append([t.clone() for t in dec(emb, [], pe, sparse, dense, True)])
        torch.cuda.synchronize()
        assert dec._workspace.data_ptr() == ws.data_ptr() and _guard_intact(ws, need)
    for o in outs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(o, outs[0])), "decoder output depends on stale workspace contents"

    vcfg = aco.SMALL_HD64
    tower = ac.VisionTransformer_MaPLe(vcfg.input_resolution, vcfg.patch_size, vcfg.width, vcfg.layers, vcfg.heads,
                                       vcfg.output_dim, design_details={"trainer": "MaPLe", "maple_length": vcfg.n_ctx}).eval()
    tower.load_state_dict(aco.seeded_vit_state(vcfg, 0), strict=True)
    tower = tower.cuda()
    shared, deep = aco.seeded_prompts(vcfg, depth=3)
    shared, deep = shared.cuda(), [d.cuda() for d in deep]
    img = torch.randn(3, 3, vcfg.input_resolution, vcfg.input_resolution, generator=gen).cuda()
    alpha = (torch.rand(3, 1, vcfg.input_resolution, vcfg.input_resolution, generator=gen) > 0.5).float().cuda()
    with torch.no_grad():
        tower(img, alpha, shared, deep)
    need = tower._workspace.numel()
    feats = []
    for fill in PATTERNS:
        ws = _workspace_with_guard(need, fill)
        tower._workspace = ws[:need]
        with torch.no_grad():
            feats.append(tower(img, alpha, shared, deep).clone())
        torch.cuda.synchronize()
        assert tower._workspace.data_ptr() == ws.data_ptr() and _guard_intact(ws, need)
    assert all(torch.equal(f, feats[0]) for f in feats[1:]), "tower output depends on stale workspace contents"
