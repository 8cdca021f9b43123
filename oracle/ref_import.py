"""Import the UNMODIFIED reference encoder from /root/reference (container only).

TEST INFRASTRUCTURE. Nothing under ``camouflaged_vlm_b200/`` may import this file.
``/root/reference`` does not exist on the GPU box, so only ``oracle/make_golden.py``
and the ``-m "not gpu"`` tests that are skipped when the tree is absent use it.

``import models`` fails in this image (mmcv / open_clip / dassl are missing), but
``models/mmseg/models/sam/image_encoder.py`` only needs torch and ``.common``
(image_encoder.py:7-22).  We therefore register an empty stub package whose
``__path__`` points at the reference directory and load the two files below it.
Nothing is copied; the sources stay where they lie.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("CVLM_REFERENCE_ROOT", "/root/reference")
SAM_DIR = os.path.join(REF_ROOT, "models", "mmseg", "models", "sam")
# the two encoder files staged by oracle/stage_ref.py (git-ignored copy that travels to the GPU box, where
# /root/reference does not exist); only the encoder is staged -- decoder / CLIP goldens are made in the container
STAGED_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "refsam")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(SAM_DIR, "image_encoder.py"))


def staged_encoder_available() -> bool:
    return all(os.path.isfile(os.path.join(STAGED_DIR, f)) for f in ("image_encoder.py", "common.py"))


def load_staged_encoder_module():
    """The reference's image_encoder module from oracle/_ref/refsam (unmodified files, see oracle/stage_ref.py)."""
    if not staged_encoder_available():
        raise FileNotFoundError(f"no staged reference under {STAGED_DIR}: run oracle/stage_ref.py in the container")
    if "refsam_staged" not in sys.modules:
        pkg = types.ModuleType("refsam_staged")
        pkg.__path__ = [STAGED_DIR]
        sys.modules["refsam_staged"] = pkg
    mods = []
    for name in ("common", "image_encoder"):
        full = f"refsam_staged.{name}"
        if full not in sys.modules:
            spec = importlib.util.spec_from_file_location(full, os.path.join(STAGED_DIR, f"{name}.py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            spec.loader.exec_module(mod)
        mods.append(sys.modules[full])
    return mods[1]


def _load(name: str):
    full = f"refsam.{name}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, os.path.join(SAM_DIR, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_modules():
    """Returns (image_encoder_module, common_module) of the reference."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")
    if "refsam" not in sys.modules:
        pkg = types.ModuleType("refsam")
        pkg.__path__ = [SAM_DIR]
        sys.modules["refsam"] = pkg
    common = _load("common")
    enc = _load("image_encoder")
    return enc, common


def load_reference_decoder_modules():
    """mask_decoder_edge.py + transformer_maskdecoder_edge.py (only need .common)."""
    load_reference_modules()
    return _load("mask_decoder_edge"), _load("transformer_maskdecoder_edge")


def build_reference_encoder(cfg, staged: bool = False):
    """Instantiate the reference ImageEncoderViT the way models/sam_maskdecoder_edge.py:119-135 does.
    ``staged=True`` takes the class from oracle/_ref/refsam (the GPU box) instead of /root/reference."""
    from functools import partial

    import torch

    enc = load_staged_encoder_module() if staged else load_reference_modules()[0]
    return enc.ImageEncoderViT(
        img_size=cfg.img_size,
        patch_size=cfg.patch_size,
        in_chans=cfg.in_chans,
        embed_dim=cfg.embed_dim,
        depth=cfg.depth,
        num_heads=cfg.num_heads,
        mlp_ratio=cfg.mlp_ratio,
        out_chans=cfg.out_chans,
        qkv_bias=cfg.qkv_bias,
        norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
        act_layer=torch.nn.GELU,
        use_abs_pos=cfg.use_abs_pos,
        use_rel_pos=cfg.use_rel_pos,
        rel_pos_zero_init=True,
        window_size=cfg.window_size,
        global_attn_indexes=tuple(cfg.global_attn_indexes),
    )


def load_reference_alpha_clip_model():
    """alpha_clip_rw/model.py of the reference (VisionTransformer_MaPLe: SURVEY.md 8 f2).  The file imports `loralib`
    (absent from this image) at module scope but only touches it when lora_adapt=True, which the cascade never sets
    (alpha_clip_rw/model.py:203-206, 263-267); an empty stand-in module lets the unmodified source load."""
    path = os.path.join(REF_ROOT, "alpha_clip_rw", "model.py")
    if not os.path.isfile(path):
        raise FileNotFoundError(f"reference Alpha-CLIP model not found at {path}")
    name = "refalphaclip_model"
    if name in sys.modules:
        return sys.modules[name]
    if "loralib" not in sys.modules:
        sys.modules["loralib"] = types.ModuleType("loralib")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
