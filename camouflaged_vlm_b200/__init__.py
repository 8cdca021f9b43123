"""B200-native SAM-Adapter image encoder (hot path of intcomp/camouflaged-vlm).

    from camouflaged_vlm_b200 import ImageEncoderViT      # drop-in for the reference class
"""
from .config import PRESETS, VIT_B, VIT_H, VIT_L, EncoderConfig, global_only  # noqa: F401
from .image_encoder import ImageEncoderViT, build_encoder  # noqa: F401

__all__ = ["ImageEncoderViT", "build_encoder", "EncoderConfig", "PRESETS", "VIT_H", "VIT_L", "VIT_B", "global_only"]
