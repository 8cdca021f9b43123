"""Pin the oracle (oracle/encoder_oracle.py) against outputs of the reference itself.

tests/golden/*.npz were produced by oracle/make_golden.py, which imports the unmodified reference
module from /root/reference.  When that tree is present (build container) the oracle is also
compared with the live reference, including the high-pass filter and padded-window semantics.
"""
import os

import numpy as np
import pytest
import torch

from camouflaged_vlm_b200.config import PRESETS
from camouflaged_vlm_b200.synth import param_shapes, seeded_images, seeded_state_dict
from oracle import encoder_oracle as orc
from oracle.ref_import import build_reference_encoder, reference_available

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FP32_TOL = 1e-5     # north_star: <= 1e-5 relative L2 on the fp32 check path


@pytest.mark.parametrize("name,batch", [("tiny", 2), ("small_hd80", 1), ("small_hd64", 1)])
def test_oracle_matches_reference_golden(name, batch):
    cfg = PRESETS[name]
    g = np.load(os.path.join(GOLDEN, f"encoder_{name}.npz"))
    out, interm = orc.encoder_forward(seeded_state_dict(cfg), cfg, seeded_images(cfg, batch), interm=True)
    assert out.shape == g["out"].shape
    assert orc.rel_l2(out, torch.from_numpy(g["out"])) < FP32_TOL
    assert len(interm) == len(cfg.global_attn_indexes)
    for i, t in enumerate(interm):
        assert orc.rel_l2(t, torch.from_numpy(g[f"interm{i}"])) < FP32_TOL


@pytest.mark.slow
def test_oracle_matches_reference_golden_vit_h():
    """BASELINE.json config 1: ViT-H, batch 1, 1024^2, fp32 on CPU (about 20-30 s)."""
    cfg = PRESETS["vit_h"]
    g = np.load(os.path.join(GOLDEN, "encoder_vit_h_sub.npz"))
    out, interm = orc.encoder_forward(seeded_state_dict(cfg), cfg, seeded_images(cfg, 1), interm=True)
    assert tuple(out.shape) == (1, 256, 64, 64)
    assert orc.rel_l2(out[:, :, ::8, ::8], torch.from_numpy(g["out"])) < FP32_TOL
    assert abs(float(out.double().norm()) - float(g["out_norm"])) / float(g["out_norm"]) < 1e-5
    for i, t in enumerate(interm):
        assert orc.rel_l2(t[:, ::16, ::8, ::8], torch.from_numpy(g[f"interm{i}"])) < FP32_TOL


def test_state_dict_names_match_reference_count():
    """527 keys / 637.213 M params for ViT-H"""
    shapes = param_shapes(PRESETS["vit_h"])
    assert len(shapes) == 527
    assert sum(1 for k in shapes if k.startswith("prompt_generator.")) == 2 + 2 + 2 * 32 + 2
    n = sum(int(np.prod(s)) for s in shapes.values())
    assert abs(n - 637.213e6) < 1e3


def test_padded_keys_take_part_in_softmax():
    """no mask in window attention; a masked variant must differ."""
    cfg = PRESETS["tiny"]
    sd = seeded_state_dict(cfg)
    x = torch.randn(1, cfg.grid, cfg.grid, cfg.embed_dim, generator=torch.Generator().manual_seed(1))
    y = orc.block(x, sd, "blocks.0.", cfg.num_heads, cfg.window_size, True, 1e-6)
    # drop the padding instead: run attention on the un-padded corner window only -> different result
    ws = cfg.window_size
    rem = cfg.grid % ws
    corner = x[:, -rem:, -rem:, :]
    yc = orc.block(torch.nn.functional.pad(corner, (0, 0, 0, ws - rem, 0, ws - rem)), sd, "blocks.0.",
                   cfg.num_heads, 0, False, 1e-6)
    assert y.shape == x.shape and yc.shape[1] == ws


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
class TestAgainstLiveReference:
    def test_high_pass_equals_reference_fft(self):
        cfg = PRESETS["tiny"]
        ref = build_reference_encoder(cfg)
        x = seeded_images(cfg, 2, seed=3)
        assert orc.rel_l2(orc.high_pass(x, 0.25), ref.prompt_generator.fft(x, 0.25)) < 1e-6

    def test_state_dict_keys_and_shapes(self):
        for name in ("tiny", "small_hd80"):
            cfg = PRESETS[name]
            ref_sd = build_reference_encoder(cfg).state_dict()
            mine = param_shapes(cfg)
            assert list(ref_sd.keys()) == list(mine.keys())
            for k, v in ref_sd.items():
                assert tuple(v.shape) == mine[k], k

    def test_forward_without_interm_and_other_seed(self):
        cfg = PRESETS["tiny"]
        ref = build_reference_encoder(cfg).eval()
        sd = seeded_state_dict(cfg, seed=5)
        ref.load_state_dict(sd, strict=True)
        x = seeded_images(cfg, 1, seed=5)
        with torch.no_grad():
            want = ref(x)
        got = orc.encoder_forward(sd, cfg, x)
        assert orc.rel_l2(got, want) < FP32_TOL

    def test_global_only_variant(self):
        from camouflaged_vlm_b200.config import global_only
        cfg = global_only(PRESETS["tiny"])
        ref = build_reference_encoder(cfg).eval()
        sd = seeded_state_dict(cfg, seed=2)
        ref.load_state_dict(sd, strict=True)
        x = seeded_images(cfg, 1, seed=2)
        with torch.no_grad():
            want = ref(x)
        assert orc.rel_l2(orc.encoder_forward(sd, cfg, x), want) < FP32_TOL


@pytest.mark.parametrize("name,batch", [("tiny", 2), ("small_hd80", 1), ("small_hd64", 1)])
def test_oracle_adapter_gradients_match_reference_autograd(name, batch):
    """the gradients of every prompt_generator.* tensor through the oracle (torch autograd over the
    functional restatement) equal the UNMODIFIED reference's under the training script's freeze rule
    (tests/golden/encoder_grads_*.npz, oracle/make_golden_grads.py)."""
    cfg = PRESETS[name]
    gold = np.load(os.path.join(GOLDEN, f"encoder_grads_{name}.npz"))
    c_out, c_int = orc.cotangents(cfg, batch, 0)
    grads = orc.adapter_gradients(seeded_state_dict(cfg, 0), cfg, seeded_images(cfg, batch, 0), c_out, c_int)
    assert set(grads) == set(k for k in gold.files if k != "loss")
    for k, v in grads.items():
        assert orc.rel_l2(v, torch.from_numpy(gold[k])) < 1e-4, k
