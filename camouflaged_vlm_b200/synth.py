"""Seeded synthetic weights and inputs (there is no network for checkpoints or datasets).

Recipe from SURVEY.md section 8d: every Linear/Conv2d weight and bias, ``pos_embed`` and
``rel_pos_h/w`` ~ N(0, 0.02); LayerNorm / LayerNorm2d weight ~ 1 + N(0, 0.02), bias ~ N(0, 0.02).
The default constructor leaves pos_embed / rel_pos at zero (image_encoder.py:78-80, 485-486),
which would make those paths vacuously correct, so parity runs always use this fill.

Values come from ``numpy.random.RandomState`` seeded per tensor NAME, so they are identical on
every platform / torch version and do not depend on parameter order.
"""
import zlib
from typing import Dict, Tuple

import numpy as np
import torch

from .config import EncoderConfig


def param_shapes(cfg: EncoderConfig) -> Dict[str, Tuple[int, ...]]:
    """state_dict names and shapes of the reference encoder (image_encoder.py:65-130), in its order."""
    d, g, r, p = cfg.embed_dim, cfg.grid, cfg.rank, cfg.patch_size
    out: Dict[str, Tuple[int, ...]] = {}
    if cfg.use_abs_pos:
        out["pos_embed"] = (1, g, g, d)
    out["patch_embed.proj.weight"] = (d, cfg.in_chans, p, p)
    out["patch_embed.proj.bias"] = (d,)
    for i in range(cfg.depth):
        b = f"blocks.{i}."
        s = g if cfg.block_window(i) == 0 else cfg.window_size
        out[b + "norm1.weight"] = (d,)
        out[b + "norm1.bias"] = (d,)
        if cfg.use_rel_pos:
            out[b + "attn.rel_pos_h"] = (2 * s - 1, cfg.head_dim)
            out[b + "attn.rel_pos_w"] = (2 * s - 1, cfg.head_dim)
        out[b + "attn.qkv.weight"] = (3 * d, d)
        if cfg.qkv_bias:
            out[b + "attn.qkv.bias"] = (3 * d,)
        out[b + "attn.proj.weight"] = (d, d)
        out[b + "attn.proj.bias"] = (d,)
        out[b + "norm2.weight"] = (d,)
        out[b + "norm2.bias"] = (d,)
        out[b + "mlp.lin1.weight"] = (cfg.mlp_dim, d)
        out[b + "mlp.lin1.bias"] = (cfg.mlp_dim,)
        out[b + "mlp.lin2.weight"] = (d, cfg.mlp_dim)
        out[b + "mlp.lin2.bias"] = (d,)
    out["neck.0.weight"] = (cfg.out_chans, d, 1, 1)
    out["neck.1.weight"] = (cfg.out_chans,)
    out["neck.1.bias"] = (cfg.out_chans,)
    out["neck.2.weight"] = (cfg.out_chans, cfg.out_chans, 3, 3)
    out["neck.3.weight"] = (cfg.out_chans,)
    out["neck.3.bias"] = (cfg.out_chans,)
    pg = "prompt_generator."
    out[pg + "shared_mlp.weight"] = (d, r)
    out[pg + "shared_mlp.bias"] = (d,)
    out[pg + "embedding_generator.weight"] = (r, d)
    out[pg + "embedding_generator.bias"] = (r,)
    for i in range(cfg.depth):
        out[pg + f"lightweight_mlp_{i}.0.weight"] = (r, r)
        out[pg + f"lightweight_mlp_{i}.0.bias"] = (r,)
    out[pg + "prompt_generator.proj.weight"] = (r, 3, p, p)
    out[pg + "prompt_generator.proj.bias"] = (r,)
    return out


_NORM_WEIGHTS = ("norm1.weight", "norm2.weight", "neck.1.weight", "neck.3.weight")


def _normal(name: str, shape, seed: int) -> np.ndarray:
    rs = np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0xFFFFFFFF)
    return rs.standard_normal(size=shape).astype(np.float32)


def seeded_state_dict(cfg: EncoderConfig, seed: int = 0, std: float = 0.02) -> Dict[str, torch.Tensor]:
    sd = {}
    for name, shape in param_shapes(cfg).items():
        v = _normal(name, shape, seed) * np.float32(std)
        if name.endswith(_NORM_WEIGHTS):
            v = v + np.float32(1.0)
        sd[name] = torch.from_numpy(v)
    return sd


def seeded_images(cfg: EncoderConfig, batch: int, seed: int = 0) -> torch.Tensor:
    """ImageNet-normalised inputs are ~N(0,1) (datasets/wrappers.py:22-27)."""
    imgs = [_normal(f"image.{seed}.{b}", (cfg.in_chans, cfg.img_size, cfg.img_size), seed)
            for b in range(batch)]
    return torch.from_numpy(np.stack(imgs, 0))
