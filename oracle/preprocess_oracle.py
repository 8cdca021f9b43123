"""CPU restatement of the reference's image input pipeline (datasets/wrappers.py:22-27, 61):

    transforms.Resize((S, S))   -> PIL Image.resize((S, S), BILINEAR)            (Pillow 8bpc antialiased resample)
    transforms.ToTensor()       -> uint8 HWC -> float32 CHW / 255
    transforms.Normalize(mean, std)

TEST INFRASTRUCTURE -- NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module; the product path is csrc/preprocess.cu behind cvlm_op_preprocess_u8.

The resample is a third-party dependency of the reference (Pillow, src/libImaging/Resample.c; `pillow` is unpinned in
the reference's requirements.txt, the container has 12.2.0).  Its published algorithm, restated here:
  * two separable passes, horizontal first, each rounding to uint8 (the vertical pass reads the rounded image);
  * per output coordinate xx: scale = in/out, filterscale = max(scale, 1), support = 1.0 * filterscale (bilinear),
    center = (xx + 0.5) * scale, xmin = int(center - support + 0.5) clipped at 0, xmax = int(center + support + 0.5)
    clipped at `in`, weights w = triangle((x + xmin - center + 0.5) / filterscale) normalised by their sum;
  * 8-bit path: weights are rounded to integers kk = int(0.5 + w * 2^22) (PRECISION_BITS = 32 - 8 - 2), the
    accumulator starts at 2^21, and the result is clip(acc >> 22, 0, 255).
Pinned bit-exactly against PIL + torchvision on this container's versions (tests/golden/preprocess_*.npz,
oracle/make_golden_preprocess.py; tests/test_preprocess_oracle.py also checks live when PIL is importable).
"""
import math
from typing import Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2
IMAGENET_MEAN = (0.485, 0.456, 0.406)          # datasets/wrappers.py:25-26
IMAGENET_STD = (0.229, 0.224, 0.225)


def _triangle(x: float) -> float:
    """bilinear_filter, Resample.c."""
    if x < 0.0:
        x = -x
    return 1.0 - x if x < 1.0 else 0.0


def precompute_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """precompute_coeffs + normalize_coeffs_8bpc (Resample.c) for the full-image box [0, in_size).
    Returns (bounds int32 [out, 2] = (xmin, count), kk int32 [out, ksize], ksize)."""
    scale = float(in_size) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_triangle((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One 8bpc pass along `axis` (0 = vertical, 1 = horizontal) of a uint8 [H, W, C] image."""
    in_size = img.shape[axis]
    bounds, kk, ksize = precompute_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)                 # [in, other, C]
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(cnt):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL Image.resize((out_w, out_h), BILINEAR) of a uint8 [H, W, C] image: horizontal pass, then vertical."""
    assert img.dtype == np.uint8 and img.ndim == 3
    h, w = img.shape[:2]
    if w != out_w:
        img = _resample_axis(img, out_w, 1)
    if h != out_h:
        img = _resample_axis(img, out_h, 0)
    return img


def preprocess_u8(img: np.ndarray, size: int, mean: Sequence[float] = IMAGENET_MEAN,
                  std: Sequence[float] = IMAGENET_STD) -> np.ndarray:
    """Resize -> ToTensor -> Normalize: uint8 [H, W, 3] -> float32 [3, size, size] (fp32 arithmetic, in
    torchvision's order: x / 255, then (x - mean) / std)."""
    r = resize_bilinear_u8(img, size, size)
    x = r.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)
    m = np.asarray(mean, np.float32)[:, None, None]
    s = np.asarray(std, np.float32)[:, None, None]
    return ((x - m) / s).astype(np.float32)
