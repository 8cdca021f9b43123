"""The CUDA image input pipeline (cvlm_op_preprocess_u8 through the C ABI) against the oracle and the golden vectors
of PIL + torchvision: integer/byte work and three IEEE fp32 ops per value, so the bar is bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from camouflaged_vlm_b200.preprocess import preprocess_u8      # noqa: E402
from oracle import preprocess_oracle as po                     # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz")


def test_preprocess_matches_golden_vectors_bit_exactly():
    g = np.load(GOLDEN)
    for i, (h, w, s) in enumerate(g["cases"]):
        got = preprocess_u8(torch.from_numpy(g[f"in{i}"]).cuda(), int(s))[0].cpu().numpy()
        assert np.array_equal(got, g[f"out{i}"]), (h, w, s)


@pytest.mark.parametrize("h,w", [(1024, 1024), (300, 400), (2000, 3000), (1024, 700), (50, 2048), (1, 1)])
def test_preprocess_1024_matches_oracle_bit_exactly(h, w):
    """BASELINE size: any input -> 1024^2 (identity, up-scale, wide down-scale kernels, one axis only, degenerate)."""
    rng = np.random.RandomState(h * 7 + w)
    img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    got = preprocess_u8(torch.from_numpy(img).cuda(), 1024)[0].cpu().numpy()
    assert np.array_equal(got, po.preprocess_u8(img, 1024))


def test_preprocess_batch_of_mixed_sizes_and_argument_checks():
    rng = np.random.RandomState(3)
    imgs = [rng.randint(0, 256, (h, w, 3)).astype(np.uint8) for h, w in [(120, 80), (64, 64), (33, 250)]]
    got = preprocess_u8([torch.from_numpy(i).cuda() for i in imgs], 64).cpu().numpy()
    assert got.shape == (3, 3, 64, 64)
    for k, im in enumerate(imgs):
        assert np.array_equal(got[k], po.preprocess_u8(im, 64))
    with pytest.raises(RuntimeError):
        preprocess_u8(torch.from_numpy(imgs[0]), 64)                       # no CPU fallback
    with pytest.raises(ValueError):
        preprocess_u8(torch.zeros(8, 8, 4, dtype=torch.uint8, device="cuda"), 64)
