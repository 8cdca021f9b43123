"""CPU-side checks of the mask decoder's host logic: the oracle's full output set against the
reference's golden vectors, the drop-in module's parameter tree, and the weight re-layouts that turn the transposed
convolutions into GEMMs (checked against torch's own conv_transpose2d in fp32)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from camouflaged_vlm_b200 import mask_decoder as md
from oracle import mask_decoder_oracle as mdo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def make_decoder():
    return md.MaskDecoder(num_multimask_outputs=3, transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256,
                          transformer=md.TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8))


def test_oracle_all_outputs_match_reference_golden():
    g = np.load(os.path.join(GOLDEN, "decoder_small.npz"))
    got = mdo.decode_all(mdo.seeded_decoder_state(0), torch.from_numpy(g["emb"]))
    for key, name in (("masks", "masks_all"), ("edge", "edge"), ("iou", "iou")):
        want = torch.from_numpy(g[name])
        assert got[key].shape == want.shape
        assert float((got[key] - want).abs().max()) < 1e-5, name


def test_module_parameter_tree_matches_reference():
    dec = make_decoder()
    shapes = mdo.decoder_param_shapes()
    sd = dec.state_dict()
    assert set(sd) == set(shapes)
    for k, v in sd.items():
        assert tuple(v.shape) == shapes[k], k
    ref = {k: v for k, v in mdo.seeded_decoder_state(0).items() if not k.startswith("probe.")}
    dec.load_state_dict(ref, strict=True)
    assert dec.num_mask_tokens == 5 and md.MaskDecoder_Edge is md.MaskDecoder


def test_module_has_no_cpu_path():
    dec = make_decoder()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):
        dec(torch.zeros(1, 256, 8, 8), [], torch.zeros(1, 256, 8, 8), torch.zeros(1, 2, 256), torch.zeros(1, 256, 8, 8), False)


def nested_index(g):
    """raster (Y, X) of the 4g x 4g map -> nested row index (csrc/decoder.cu raster_to_nested)."""
    Y, X = torch.meshgrid(torch.arange(4 * g), torch.arange(4 * g), indexing="ij")
    return ((((Y >> 2) * g + (X >> 2)) * 4 + ((Y >> 1) & 1) * 2 + ((X >> 1) & 1)) << 2) + (Y & 1) * 2 + (X & 1)


def test_transposed_convolutions_as_gemms():
    torch.manual_seed(0)
    g = 8
    rb = lambda t: t.to(torch.bfloat16).float()              # weights exactly representable in the packed dtype
    sd = {k: (rb(v) if v.dim() == 4 else v) for k, v in mdo.seeded_decoder_state(3).items() if not k.startswith("probe.")}
    packed = md.pack_decoder_weights(sd, 2, torch.device("cpu"))
    x = torch.randn(1, 256, g, g)
    # two k2 s2 transposed convolutions == two GEMMs whose outputs are reshaped (nested pixel order)
    rows = x[0].flatten(1).t()                                                        # [g*g, 256]
    u1 = rows @ packed["output_upscaling.0.wp"].float().t() + packed["output_upscaling.0.bp"]
    h1 = u1.reshape(-1, 64)                                                           # [4 g*g, 64]
    u2 = h1 @ packed["output_upscaling.3.wp"].float().t() + packed["output_upscaling.3.bp"]
    nested = u2.reshape(-1, 32)                                                       # [16 g*g, 32]
    want = F.conv_transpose2d(F.conv_transpose2d(x, sd["output_upscaling.0.weight"], sd["output_upscaling.0.bias"], stride=2),
                              sd["output_upscaling.3.weight"], sd["output_upscaling.3.bias"], stride=2)
    got = nested[nested_index(g).flatten()].t().reshape(1, 32, 4 * g, 4 * g)
    assert float((got - want).abs().max()) < 1e-4
    # 3x3 s1 p1 transposed convolution == im2col GEMM with the flipped kernel
    y = torch.randn(1, 32, 4 * g, 4 * g)
    cols = F.unfold(y, 3, padding=1)[0]                                               # [32*9, P] ordered (c, ty, tx)
    cols = cols.reshape(32, 9, -1).permute(2, 1, 0).reshape(-1, 288)                  # [P, (ty*3+tx)*32 + c]
    wp = packed["embedding_maskfeature.0.wp"].float().reshape(64, 9, 64)             # 64 channels per tap, zero padded
    assert float(wp[:, :, 32:].abs().max()) == 0.0
    got = (cols @ wp[:, :, :32].reshape(64, 288).t() + sd["embedding_maskfeature.0.bias"]).t().reshape(1, 64, 4 * g, 4 * g)
    want = F.conv_transpose2d(y, sd["embedding_maskfeature.0.weight"], sd["embedding_maskfeature.0.bias"], stride=1, padding=1)
    assert float((got - want).abs().max()) < 1e-4
    assert packed["embedding_maskfeature.3.wp"].shape == (32, 576)
    assert packed["tokens"].shape == (6, 256)


def test_decoder_handle_workspace_and_config_validation():
    import ctypes as C
    from camouflaged_vlm_b200 import _lib
    lib = _lib.load()

    def cfg(**kw):
        c = _lib.DecoderConfigC()
        c.dim, c.heads, c.mlp_dim, c.depth, c.num_mask_tokens, c.grid, c.cond_tokens = 256, 8, 2048, 2, 4, 64, 2
        c.ln_eps, c.ln2d_eps = 1e-5, 1e-6
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    h = C.c_void_p()
    assert lib.cvlm_decoder_create(C.byref(cfg()), C.byref(h)) == 0
    per_image = [lib.cvlm_decoder_workspace_bytes(h, b) / b for b in (1, 8, 32)]
    assert all(80e6 < x < 100e6 for x in per_image), per_image          # no im2col matrix: < 100 MB per image
    assert lib.cvlm_decoder_workspace_bytes(h, 0) < 0
    assert lib.cvlm_decoder_destroy(h) == 0
    for bad in (dict(dim=128), dict(heads=4), dict(grid=12), dict(grid=0), dict(cond_tokens=9), dict(mlp_dim=100), dict(depth=0)):
        h2 = C.c_void_p()
        assert lib.cvlm_decoder_create(C.byref(cfg(**bad)), C.byref(h2)) == -3, bad
        assert lib.cvlm_last_error()


def test_decoder_accepts_the_reference_transformer_instance():
    from oracle.ref_import import reference_available
    if not reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    from oracle.ref_import import load_reference_decoder_modules
    _, tr = load_reference_decoder_modules()
    dec = md.MaskDecoder(num_multimask_outputs=3, transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256,
                         transformer=tr.TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8))
    assert set(dec.state_dict()) == set(mdo.decoder_param_shapes())
    packed = md.pack_decoder_weights(dict(dec.state_dict()), 2, torch.device("cpu"))
    assert packed["transformer.layers.1.mlp.lin1.weight.t"].shape == (256, 2048)
    assert packed["transformer.final_attn_token_to_image.k_proj.weight.bf16"].dtype == torch.bfloat16
