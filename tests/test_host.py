"""CPU-side checks: the C-ABI library builds, loads and exports every symbol include/cvlm.h declares
(no compute calls without a GPU), the module keeps the reference's constructor / state_dict contract,
and the host-side argument validation of the library works."""
import ctypes as C
import os
import re

import pytest
import torch

from camouflaged_vlm_b200 import PRESETS, ImageEncoderViT, _lib, build_encoder
from camouflaged_vlm_b200.packing import pack_weights, rank_pad
from camouflaged_vlm_b200.synth import param_shapes, seeded_images, seeded_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "cvlm.h")).read()
    declared = set(re.findall(r"\b(cvlm_[a-z0-9_]+)\s*\(", header))
    declared.discard("cvlm_encoder")     # the opaque struct
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cvlm_abi_version() == 1


def test_host_side_validation_without_gpu(lib):
    cc = _lib.EncoderConfigC()
    h = C.c_void_p()
    assert lib.cvlm_encoder_create(C.byref(cc), C.byref(h)) < 0           # zeroed config
    assert b"patch_size" in lib.cvlm_last_error()
    cfg = PRESETS["vit_h"]
    cc.img_size, cc.patch_size, cc.in_chans, cc.embed_dim, cc.depth = 1024, 16, 3, 1280, 32
    cc.num_heads, cc.mlp_dim, cc.out_chans, cc.window_size = 16, 5120, 256, 14
    cc.rank, cc.rank_pad, cc.use_abs_pos = 40, 40, 1
    assert lib.cvlm_encoder_create(C.byref(cc), C.byref(h)) == 0
    per_image = lib.cvlm_encoder_workspace_bytes(h, 1)
    assert 115e6 < per_image < 135e6                                      # ~127 MB / image
    assert abs(lib.cvlm_encoder_workspace_bytes(h, 32) - 32 * per_image) < 1e5
    assert lib.cvlm_encoder_set_tensor(h, b"patch.w", None, 16) < 0       # null pointer rejected
    assert lib.cvlm_encoder_destroy(h) == 0
    cc.num_heads = 10                                                     # head_dim 128: unsupported
    assert lib.cvlm_encoder_create(C.byref(cc), C.byref(h)) == -3


def test_module_contract_matches_reference():
    cfg = PRESETS["vit_h"]
    with torch.device("meta"):
        m = build_encoder(cfg)
    sd = m.state_dict()
    want = param_shapes(cfg)
    assert list(sd.keys()) == list(want.keys())                           # 527 names, reference order
    for k, v in sd.items():
        assert tuple(v.shape) == want[k] and v.dtype == torch.float32, k
    assert m.img_size == 1024 and m.blocks[7].window_size == 0 and m.blocks[0].window_size == 14
    assert m.num_stages == 32 and m.out_indices == tuple(range(32)) and m.scale_factor == 32
    trainable = [n for n, _ in m.named_parameters() if "prompt_generator" in n]
    assert len(trainable) == 70                                           # train_...py:301-303 filter


def test_module_rejects_unsupported_layers_and_cpu_forward():
    with pytest.raises(NotImplementedError):
        ImageEncoderViT(norm_layer=torch.nn.BatchNorm1d)
    with pytest.raises(NotImplementedError):
        ImageEncoderViT(act_layer=torch.nn.ReLU)
    cfg = PRESETS["tiny"]
    m = build_encoder(cfg).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(seeded_images(cfg, 1))


def test_packing_shapes_and_folds():
    cfg = PRESETS["small_hd80"]
    sd = seeded_state_dict(cfg)
    p = pack_weights(sd, cfg, "cpu", dtype=torch.bfloat16)
    rp = rank_pad(cfg.rank)
    assert p["lw.w"].shape == (cfg.depth * rp, rp) and p["lw.w"].dtype == torch.bfloat16
    assert p["emb.cst"].shape == (cfg.tokens, rp) and p["emb.cst"].dtype == torch.float32
    d = cfg.embed_dim
    # K / V carry no Linear bias (folded away); what is left on them is W beta of the folded norm1
    w, beta = sd["blocks.0.attn.qkv.weight"].double(), sd["blocks.0.norm1.bias"].double()
    want_b = w @ beta
    want_b[:d] += sd["blocks.0.attn.qkv.bias"][:d].double()
    assert torch.allclose(p["blk.0.qkv.b"].double(), want_b, atol=1e-6)
    # norm1 folded into qkv: gamma scales the weight columns, colsum is taken over the ROUNDED operand
    gamma = sd["blocks.0.norm1.weight"]
    assert torch.equal(p["blk.0.qkv.w"], (sd["blocks.0.attn.qkv.weight"] * gamma[None, :]).bfloat16())
    assert torch.allclose(p["blk.0.qkv.c"].double(), p["blk.0.qkv.w"].double().sum(1), atol=1e-5)
    b_v = sd["blocks.0.attn.qkv.bias"][2 * d:]
    want = sd["blocks.0.attn.proj.bias"] + sd["blocks.0.attn.proj.weight"] @ b_v
    assert torch.allclose(p["blk.0.proj.b"], want, atol=1e-6)
    assert p["blk.0.relh"].shape == (2 * 14 - 1, 80) and p["blk.2.relh"].shape == (2 * cfg.grid - 1, 80)
    for t in p.values():
        assert t.is_contiguous()


def test_options_round_trip_without_gpu(lib):
    """cvlm_set_option / cvlm_get_option: the launchers read these instead of the environment (no getenv per launch)."""
    for name, value in (("gemm_mode", 2), ("dec_cluster", 4), ("attn_legacy", 1), ("attn_poly", 6), ("win_mode", 1)):
        before = _lib.get_option(name)
        _lib.set_option(name, value)
        assert _lib.get_option(name) == value
        _lib.set_option(name, before)
    with pytest.raises(_lib.CvlmError):
        _lib.set_option("no_such_option", 1)


def test_staged_reference_is_the_reference_verbatim():
    """oracle/stage_ref.py: oracle/_ref/refsam holds byte-identical copies of the reference's two encoder files and the
    staged module computes what the in-place reference computes (container only)."""
    import filecmp
    from oracle import ref_import, stage_ref
    if not ref_import.reference_available():
        pytest.skip("needs /root/reference")
    assert stage_ref.stage(verbose=False)
    for f in stage_ref.FILES:
        assert filecmp.cmp(os.path.join(stage_ref.SRC_DIR, f), os.path.join(stage_ref.DST_DIR, f), shallow=False)
    cfg = PRESETS["tiny"]
    sd = seeded_state_dict(cfg, 0)
    a, b = ref_import.build_reference_encoder(cfg).eval(), ref_import.build_reference_encoder(cfg, staged=True).eval()
    a.load_state_dict(sd, strict=True)
    b.load_state_dict(sd, strict=True)
    x = seeded_images(cfg, 1)
    with torch.no_grad():
        assert torch.equal(a(x), b(x))
