"""The image-input oracle (oracle/preprocess_oracle.py) against the golden vectors produced by the reference's own
third-party implementation (PIL resize + torchvision ToTensor / Normalize, oracle/make_golden_preprocess.py) --
bit-exact, bytes and floats -- and against PIL live when it is importable."""
import os

import numpy as np
import pytest

from oracle import preprocess_oracle as po

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz")


def test_oracle_matches_golden_vectors_bit_exactly():
    g = np.load(GOLDEN)
    for i, (h, w, s) in enumerate(g["cases"]):
        img = g[f"in{i}"]
        assert img.shape == (h, w, 3)
        assert np.array_equal(po.resize_bilinear_u8(img, int(s), int(s)), g[f"u8_{i}"]), (h, w, s)
        got = po.preprocess_u8(img, int(s))
        assert got.dtype == np.float32 and np.array_equal(got, g[f"out{i}"]), (h, w, s)


def test_coefficient_tables_are_normalised_and_bounded():
    for n_in, n_out in [(53, 64), (333, 96), (64, 64), (3000, 1024), (5, 1024)]:
        bounds, kk, ksize = po.precompute_coeffs(n_in, n_out)
        assert bounds[:, 0].min() >= 0 and (bounds[:, 0] + bounds[:, 1]).max() <= n_in and bounds[:, 1].max() <= ksize
        sums = np.array([kk[i, :bounds[i, 1]].sum() for i in range(n_out)])
        assert np.abs(sums - (1 << po.PRECISION_BITS)).max() <= ksize          # rounding of each tap only


def test_oracle_matches_pil_live():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.RandomState(7)
    for h, w, s in [(41, 29, 32), (200, 310, 128), (128, 128, 128)]:
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((s, s), Image.BILINEAR))
        assert np.array_equal(po.resize_bilinear_u8(img, s, s), ref)
