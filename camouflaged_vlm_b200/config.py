"""Encoder shapes.

The constructor arguments mirror the reference ``ImageEncoderViT.__init__``
(models/mmseg/models/sam/image_encoder.py:26-44); the presets are the values the
callers pass (configs/demo.yaml:4-31, models/sam_maskdecoder_edge.py:119-135) and
the ViT-B/L/H size table (ext/meta/sam_meta.py:1-35).
"""
from dataclasses import dataclass, field, replace
from typing import Dict, Tuple


@dataclass(frozen=True)
class EncoderConfig:
    img_size: int = 1024
    patch_size: int = 16
    in_chans: int = 3
    embed_dim: int = 1280
    depth: int = 32
    num_heads: int = 16
    mlp_ratio: float = 4.0
    out_chans: int = 256
    qkv_bias: bool = True
    use_abs_pos: bool = True
    use_rel_pos: bool = True
    window_size: int = 14
    global_attn_indexes: Tuple[int, ...] = (7, 15, 23, 31)
    # hard-coded by the reference, image_encoder.py:116-123
    scale_factor: int = 32
    freq_nums: float = 0.25

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def tokens(self) -> int:
        return self.grid * self.grid

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads

    @property
    def mlp_dim(self) -> int:
        return int(self.embed_dim * self.mlp_ratio)

    @property
    def rank(self) -> int:
        """Adapter width, embed_dim // scale_factor (image_encoder.py:252-253)."""
        return self.embed_dim // self.scale_factor

    def block_window(self, i: int) -> int:
        """image_encoder.py:93: global blocks get window_size 0."""
        return 0 if i in self.global_attn_indexes else self.window_size

    def flops_per_image(self) -> float:
        """Algorithmic FLOPs of one forward (SURVEY.md section 8d closed form): real tokens only,
        real queries x window keys for windowed blocks, 1 MAC = 2 FLOP."""
        n, d, r = self.tokens, self.embed_dim, self.rank
        hd, heads, g = self.head_dim, self.num_heads, self.grid
        lin = self.depth * (2 * n * d * 3 * d + 2 * n * d * d + 2 * 2 * n * d * self.mlp_dim)
        attn = 0.0
        for i in range(self.depth):
            w = self.block_window(i)
            if w == 0:
                attn += heads * (4 * n * n * hd + 2 * n * (2 * g) * hd)
            else:
                attn += heads * (4 * n * w * w * hd + 2 * n * (2 * w) * hd)
        k_patch = self.in_chans * self.patch_size ** 2
        patch = 2 * n * k_patch * d
        neck = 2 * n * d * self.out_chans + 2 * n * self.out_chans * self.out_chans * 9
        adapter = 2 * n * k_patch * r + 2 * n * d * r + self.depth * (2 * n * r * r + 2 * n * r * d)
        return float(lin + attn + patch + neck + adapter)


VIT_H = EncoderConfig()
VIT_L = EncoderConfig(embed_dim=1024, depth=24, num_heads=16, global_attn_indexes=(5, 11, 17, 23))
VIT_B = EncoderConfig(embed_dim=768, depth=12, num_heads=12, global_attn_indexes=(2, 5, 8, 11))

# Small shapes for CPU-speed parity tests.  They keep every structural feature of ViT-H:
# padded windows (grid % window != 0), both attention kinds, rel-pos, abs-pos, adapter.
TINY = EncoderConfig(img_size=160, embed_dim=64, depth=4, num_heads=2, out_chans=16,
                     window_size=4, global_attn_indexes=(1, 3))
# head_dim 80 / rank 10 like ViT-H, grid 18 with window 14 -> 28 padded (2x2 windows)
SMALL_HD80 = EncoderConfig(img_size=288, embed_dim=320, depth=3, num_heads=4, out_chans=256,
                           window_size=14, global_attn_indexes=(2,))
# head_dim 64 family (ViT-B/L)
SMALL_HD64 = EncoderConfig(img_size=256, embed_dim=256, depth=2, num_heads=4, out_chans=256,
                           window_size=14, global_attn_indexes=(1,))

PRESETS: Dict[str, EncoderConfig] = {
    "vit_h": VIT_H, "vit_l": VIT_L, "vit_b": VIT_B,
    "tiny": TINY, "small_hd80": SMALL_HD80, "small_hd64": SMALL_HD64,
}


def global_only(cfg: EncoderConfig) -> EncoderConfig:
    """BASELINE.json config 5: window_size=0 sweep (every block global)."""
    return replace(cfg, window_size=0, global_attn_indexes=tuple(range(cfg.depth)))
