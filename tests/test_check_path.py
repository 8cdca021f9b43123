"""fp32 check path (camouflaged_vlm_b200/check_path.py = the CUDA launch schedule in torch fp32)
against the reference's golden outputs.  north_star tolerance for this path: rel-L2 <= 1e-5."""
import os

import numpy as np
import pytest
import torch

from camouflaged_vlm_b200.check_path import encoder_forward_check, high_pass
from camouflaged_vlm_b200.config import PRESETS, global_only
from camouflaged_vlm_b200.packing import pack_weights
from camouflaged_vlm_b200.synth import seeded_images, seeded_state_dict
from oracle import encoder_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FP32_TOL = 1e-5


@pytest.mark.parametrize("name,batch", [("tiny", 2), ("small_hd80", 1), ("small_hd64", 1)])
def test_check_path_matches_reference_golden(name, batch):
    cfg = PRESETS[name]
    g = np.load(os.path.join(GOLDEN, f"encoder_{name}.npz"))
    packed = pack_weights(seeded_state_dict(cfg), cfg, "cpu", dtype=torch.float32)
    out, interm = encoder_forward_check(packed, cfg, seeded_images(cfg, batch), interm=True)
    assert orc.rel_l2(out, torch.from_numpy(g["out"])) < FP32_TOL
    for i, t in enumerate(interm):
        assert orc.rel_l2(t, torch.from_numpy(g[f"interm{i}"])) < FP32_TOL


def test_check_path_fold_equals_no_fold():
    """interm=True disables the lin2 prompt fold at global blocks; both schedules must agree."""
    cfg = PRESETS["tiny"]
    packed = pack_weights(seeded_state_dict(cfg, 4), cfg, "cpu", dtype=torch.float32)
    x = seeded_images(cfg, 1, 4)
    a = encoder_forward_check(packed, cfg, x, interm=False)
    b, _ = encoder_forward_check(packed, cfg, x, interm=True)
    assert orc.rel_l2(a, b) < 1e-6


def test_check_path_partial_matches_oracle_blocks():
    cfg = PRESETS["tiny"]
    sd = seeded_state_dict(cfg, 1)
    packed = pack_weights(sd, cfg, "cpu", dtype=torch.float32)
    x = seeded_images(cfg, 1, 1)
    for blk in range(cfg.depth):
        mine = encoder_forward_check(packed, cfg, x, stop_after_block=blk)
        want = orc.encoder_forward(sd, cfg, x, stop_after_block=blk).reshape(-1, cfg.embed_dim)
        assert orc.rel_l2(mine, want) < FP32_TOL, blk


def test_check_path_global_only_and_window_larger_than_grid():
    base = PRESETS["tiny"]
    from dataclasses import replace
    for cfg in (global_only(base), replace(base, window_size=14)):   # 14 > grid 10: one padded window
        sd = seeded_state_dict(cfg, 2)
        packed = pack_weights(sd, cfg, "cpu", dtype=torch.float32)
        x = seeded_images(cfg, 1, 2)
        assert orc.rel_l2(encoder_forward_check(packed, cfg, x), orc.encoder_forward(sd, cfg, x)) < FP32_TOL


def test_high_pass_matches_oracle():
    x = seeded_images(PRESETS["tiny"], 2, 7)
    assert orc.rel_l2(high_pass(x, 0.25), orc.high_pass(x, 0.25)) < 1e-6


def test_high_pass_projector_matrices_reproduce_the_fft():
    """csrc/highpass.cu replaces the FFT by  |x - Pr x Pr^T + Pi x Pi^T|  (packing.high_pass_matrices): exact in
    fp64 for sizes with and without 64-alignment, including the asymmetric -line frequency (Pi != 0)."""
    from camouflaged_vlm_b200.packing import high_pass_matrices
    for n in (64, 160, 288):
        x = torch.randn(1, 2, n, n, dtype=torch.float64, generator=torch.Generator().manual_seed(n))
        pcat, npr, pi = high_pass_matrices(n, 0.25, "cpu", torch.float64)
        npad = (n + 63) // 64 * 64
        assert pcat.shape == (2 * n, n) and npr.shape == (n, npad) and pi.shape == (n, npad)
        assert torch.all(npr[:, n:] == 0) and torch.all(pi[:, n:] == 0) and pi.abs().max() > 0
        pr, pim = pcat[:n], pcat[n:]
        assert torch.equal(npr[:, :n], -pr) and torch.equal(pi[:, :n], pim)
        got = (x - pr @ x @ pr.t() + pim @ x @ pim.t()).abs()
        assert orc.rel_l2(got.float(), orc.high_pass(x.float(), 0.25)) < 1e-6
