"""CPU restatement of the Alpha-CLIP visual tower with MaPLe deep prompts (the next row after the mask
decoder): VisionTransformer_MaPLe.forward (alpha_clip_rw/model.py:507-563) over CustomTransformer (:478-504),
CustomResidualAttentionBlock_MaPLe (:285-362), Attention (:176-256), LayerNorm / QuickGELU (:162-173); called by the
cascade as ``self.image_encoder(images, masks, shared_ctx, deep_compound_prompts_vision)``
(cocotrainers/mapleAlphaCLIP.py:272, 286).

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  No CUDA path exists for this row yet; this oracle is pinned
ahead of it against the reference module's own outputs (tests/golden/alpha_clip_small.npz, oracle/make_golden_alpha_clip.py,
and the live module when /root/reference is present).  Functional form over a state_dict with the reference's keys.
"""
import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from camouflaged_vlm_b200.synth import _normal

Tensor = torch.Tensor


@dataclass(frozen=True)
class VitConfig:
    input_resolution: int = 336      # ViT-L/14@336 (alpha_clip_rw/model.py build_model; mapleAlphaCLIP.py:36-48)
    patch_size: int = 14
    width: int = 1024
    layers: int = 24
    heads: int = 16
    output_dim: int = 768
    n_ctx: int = 2                   # MaPLe context tokens appended to the sequence (design_details['maple_length'])

    @property
    def grid(self) -> int:
        return self.input_resolution // self.patch_size

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + 1 + self.n_ctx


SMALL = VitConfig(input_resolution=56, patch_size=14, width=64, layers=3, heads=4, output_dim=32, n_ctx=2)
SMALL_HD64 = VitConfig(input_resolution=70, patch_size=14, width=128, layers=4, heads=2, output_dim=32, n_ctx=2)   # CUDA-path fixture


def vit_param_shapes(cfg: VitConfig) -> Dict[str, Tuple[int, ...]]:
    w, p = cfg.width, cfg.patch_size
    s: Dict[str, Tuple[int, ...]] = {
        "conv1.weight": (w, 3, p, p), "conv1_alpha.weight": (w, 1, p, p), "class_embedding": (w,),
        "positional_embedding": (cfg.grid * cfg.grid + 1, w), "ln_pre.weight": (w,), "ln_pre.bias": (w,),
        "ln_post.weight": (w,), "ln_post.bias": (w,), "proj": (w, cfg.output_dim),
    }
    for i in range(cfg.layers):
        b = f"transformer.resblocks.{i}."
        s[b + "attn.in_proj.weight"] = (3 * w, w)
        s[b + "attn.in_proj.bias"] = (3 * w,)
        s[b + "attn.out_proj.weight"] = (w, w)
        s[b + "attn.out_proj.bias"] = (w,)
        for n in ("ln_1", "ln_2"):
            s[b + n + ".weight"] = (w,)
            s[b + n + ".bias"] = (w,)
        s[b + "mlp.c_fc.weight"] = (4 * w, w)
        s[b + "mlp.c_fc.bias"] = (4 * w,)
        s[b + "mlp.c_proj.weight"] = (w, 4 * w)
        s[b + "mlp.c_proj.bias"] = (w,)
    return s


def seeded_vit_state(cfg: VitConfig, seed: int = 0) -> Dict[str, Tensor]:
    sd = {}
    for name, shape in vit_param_shapes(cfg).items():
        std = 0.5 if name.startswith("conv1") else cfg.width ** -0.5
        v = _normal("alphaclip." + name, shape, seed) * np.float32(std)
        if name.endswith(".weight") and len(shape) == 1:
            v = v * np.float32(0.1) + np.float32(1.0)
        sd[name] = torch.from_numpy(v)
    return sd


def seeded_prompts(cfg: VitConfig, depth: int, seed: int = 0) -> Tuple[Tensor, List[Tensor]]:
    """shared_ctx [n_ctx, width] and `depth - 1` deeper visual prompts (MultiModalPromptLearner, mapleAlphaCLIP.py:96-180)."""
    shared = torch.from_numpy(_normal("alphaclip.shared_ctx", (cfg.n_ctx, cfg.width), seed) * np.float32(0.02))
    deep = [torch.from_numpy(_normal(f"alphaclip.deep{i}", (cfg.n_ctx, cfg.width), seed) * np.float32(0.02))
            for i in range(depth - 1)]
    return shared, deep


def _ln(sd, name, x):
    return F.layer_norm(x.float(), (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def _attention(sd, pre, x, heads):
    """Attention.forward without mask, cosine scaling or head scaling (the defaults), x = [L, N, C]."""
    L, N, C = x.shape
    q, k, v = F.linear(x, sd[pre + "in_proj.weight"], sd[pre + "in_proj.bias"]).chunk(3, dim=-1)
    split = lambda t: t.contiguous().view(L, N * heads, -1).transpose(0, 1)
    q, k, v = split(q) * (C // heads) ** -0.5, split(k), split(v)
    a = torch.bmm(q, k.transpose(-2, -1)).softmax(dim=-1)
    o = torch.bmm(a, v).transpose(0, 1).reshape(L, N, C)
    return F.linear(o, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])


@torch.no_grad()
def vit_maple_forward(sd: Dict[str, Tensor], cfg: VitConfig, image: Tensor, alpha: Tensor, shared_ctx: Tensor,
                      deep_prompts: List[Tensor]) -> Tensor:
    """image [B,3,R,R], alpha [B,1,R,R] -> features [B, output_dim]."""
    x = F.conv2d(image.float(), sd["conv1.weight"], stride=cfg.patch_size) + \
        F.conv2d(alpha.float(), sd["conv1_alpha.weight"], stride=cfg.patch_size)
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
    x = torch.cat([sd["class_embedding"].expand(x.shape[0], 1, -1), x], dim=1) + sd["positional_embedding"]
    x = torch.cat([x, shared_ctx.expand(x.shape[0], -1, -1)], dim=1)            # VPT_shallow (:536-538)
    x = _ln(sd, "ln_pre", x).permute(1, 0, 2)                                    # NLD -> LND
    counter = 0
    for i in range(cfg.layers):
        b = f"transformer.resblocks.{i}."
        if i > 0 and len(deep_prompts) > 0 and counter <= len(deep_prompts) - 1:  # :318-331
            ctx = deep_prompts[counter].expand(x.shape[1], -1, -1).permute(1, 0, 2)
            x = torch.cat([x[:x.shape[0] - cfg.n_ctx], ctx], dim=0)
            counter += 1
        x = x + _attention(sd, b + "attn.", _ln(sd, b + "ln_1", x), cfg.heads)
        h = F.linear(_ln(sd, b + "ln_2", x), sd[b + "mlp.c_fc.weight"], sd[b + "mlp.c_fc.bias"])
        x = x + F.linear(h * torch.sigmoid(1.702 * h), sd[b + "mlp.c_proj.weight"], sd[b + "mlp.c_proj.bias"])
    x = _ln(sd, "ln_post", x.permute(1, 0, 2)[:, 0, :])
    return x @ sd["proj"]


def flops_per_image(cfg: VitConfig) -> float:
    L, w = cfg.tokens, cfg.width
    per_layer = 2.0 * L * w * (3 * w + w + 8 * w) + 4.0 * L * L * w
    return cfg.layers * per_layer + 2.0 * cfg.grid ** 2 * w * 4 * cfg.patch_size ** 2
