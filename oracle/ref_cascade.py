"""BASELINE.json configs[3]: the reference's OWN inference cascade -- ``SAM.infer_test`` (models/sam_maskdecoder_edge.py:331-357)
followed by the demo's second CLIP pass (demo.py:116-122) -- built from the unmodified reference sources, so that the
three classes this repository replaces (``ImageEncoderViT``, ``MaskDecoder_Edge``, ``VisionTransformer_MaPLe``) can be
swapped INSIDE the reference wrapper and compared: final masks (IoU) and predicted class id.

TEST INFRASTRUCTURE (same rules as the rest of oracle/).  The sources are imported from ``/root/reference`` when it exists
(this container) and otherwise from the staged copy ``oracle/_ref/cascade`` (oracle/stage_ref.py; the GPU box).  What the
image lacks is stubbed, exactly as SURVEY.md 8c prescribes:

* ``dassl`` (trainer framework: only ``TrainerX`` as a base class and ``load_checkpoint`` at import time), ``loralib``
  (touched only when ``lora_adapt=True``, never in the cascade), ``ftfy`` (tokenizer text clean-up; identity here),
  ``utils.log`` (the reference's ``utils.py`` imports tensorboardX), and the package ``__init__`` files of ``models`` /
  ``models.mmseg`` (they import mmcv and every other model family): the packages are registered as path-only stubs so
  that the files of the cascade load unmodified.
* the CLIP checkpoint download (``load_clip_to_cpu``, cocotrainers/mapleAlphaCLIP.py:27-53) is replaced by the
  reference's own ``build_model`` on a randomly initialised ViT-L/14@336 + 12-layer text tower state_dict.
Everything else -- wrapper, prompt learner, text encoder, tokenizer + vocabulary, the shipped class-name lists and text
features -- is the reference's.
"""
import contextlib
import importlib
import importlib.util
import os
import sys
import types
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("CVLM_REFERENCE_ROOT", "/root/reference")
STAGED_ROOT = os.path.join(HERE, "_ref", "cascade")
# every reference file the cascade imports or reads (oracle/stage_ref.py copies exactly these)
CASCADE_FILES = (
    "models/models.py", "models/sam_maskdecoder_edge.py", "models/ovcamo_loss.py", "models/iou_loss.py",
    "models/mmseg/models/sam/__init__.py", "models/mmseg/models/sam/sam.py", "models/mmseg/models/sam/common.py",
    "models/mmseg/models/sam/image_encoder.py", "models/mmseg/models/sam/mask_decoder.py",
    "models/mmseg/models/sam/mask_decoder_edge.py", "models/mmseg/models/sam/prompt_encoder.py",
    "models/mmseg/models/sam/transformer.py", "models/mmseg/models/sam/transformer_maskdecoder_edge.py",
    "alpha_clip_rw/__init__.py", "alpha_clip_rw/alpha_clip.py", "alpha_clip_rw/model.py", "alpha_clip_rw/simple_tokenizer.py",
    "alpha_clip_rw/bpe_simple_vocab_16e6.txt.gz", "cocotrainers/mapleAlphaCLIP.py",
    "datasets/ovcamo_info/class_names.py", "datasets/ovcamo_info/TestCamoPromptsTextFeaturesViTB-14-336.pth",
    "datasets/ovcamo_info/TrainCamoPromptsTextFeaturesViTB-14-336.pth", "configs/demo.yaml",
)


def cascade_root() -> str:
    """Where the cascade sources are: the reference tree, else the staged copy; '' when neither is present."""
    for root in (REF_ROOT, STAGED_ROOT):
        if all(os.path.isfile(os.path.join(root, f)) for f in CASCADE_FILES):
            return root
    return ""


class DotDict:
    """demo.py:21-29 (the config object the demo hands to TestMaPLeAlphaCLIP)."""

    def __init__(self, dictionary):
        for key, value in dictionary.items():
            setattr(self, key, DotDict(value) if isinstance(value, dict) else value)

    def __getitem__(self, item):
        return getattr(self, item)


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _path_package(name: str, path: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    sys.modules[name] = mod
    return mod


def _importable(name: str) -> bool:
    if name in sys.modules:          # possibly another harness's stand-in (oracle/ref_import.py registers an empty loralib)
        return True
    try:
        return importlib.util.find_spec(name) is not None
    except (ValueError, ImportError):
        return False


@contextlib.contextmanager
def _reference_imports(root: str):
    """sys.modules / sys.path / cwd set up so that the cascade files import; restored afterwards (the loaded modules stay
    alive through the objects built from them)."""
    touched = ("dassl", "dassl.engine", "dassl.utils", "dassl.optim", "loralib", "ftfy", "utils", "models", "models.models",
               "models.mmseg", "models.mmseg.models", "cocotrainers", "datasets", "datasets.ovcamo_info")
    saved = {k: sys.modules.get(k) for k in touched}
    cwd = os.getcwd()
    try:
        class TrainerX:      # base class of the (unused) trainer in cocotrainers/mapleAlphaCLIP.py
            pass

        def _no_checkpoint(*a, **k):
            raise RuntimeError("dassl stub: no checkpoints in this harness")

        _stub("dassl")
        _stub("dassl.engine", TrainerX=TrainerX)
        _stub("dassl.utils", load_checkpoint=_no_checkpoint, load_pretrained_weights=_no_checkpoint)
        _stub("dassl.optim", build_optimizer=_no_checkpoint, build_lr_scheduler=_no_checkpoint)
        if not _importable("loralib"):
            _stub("loralib")
        if not _importable("ftfy"):
            _stub("ftfy", fix_text=lambda s: s)
        _stub("utils", log=lambda *a, **k: None)
        models = _path_package("models", os.path.join(root, "models"))
        spec = importlib.util.spec_from_file_location("models.models", os.path.join(root, "models", "models.py"))
        reg = importlib.util.module_from_spec(spec)
        sys.modules["models.models"] = reg
        spec.loader.exec_module(reg)
        models.register, models.make = reg.register, reg.make
        _path_package("models.mmseg", os.path.join(root, "models", "mmseg"))
        _path_package("models.mmseg.models", os.path.join(root, "models", "mmseg", "models"))
        _path_package("cocotrainers", os.path.join(root, "cocotrainers"))
        _path_package("datasets", os.path.join(root, "datasets"))
        _path_package("datasets.ovcamo_info", os.path.join(root, "datasets", "ovcamo_info"))
        sys.path.insert(0, root)
        os.chdir(root)          # SAM.__init__ torch.load()s "./datasets/ovcamo_info/..." (sam_maskdecoder_edge.py:177-182)
        yield
    finally:
        os.chdir(cwd)
        if root in sys.path:
            sys.path.remove(root)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def random_clip_state_dict(clip_module, design_details, seed: int) -> Dict[str, torch.Tensor]:
    """A ViT-L/14@336 CLIP (visual 24 x 1024, text 12 x 768, embed 768) with the module's own random initialisation: what
    ``build_model`` sniffs from a checkpoint (alpha_clip_rw/model.py:825-848)."""
    torch.manual_seed(seed)
    clip = clip_module.CLIP(768, 336, 24, 1024, 14, 77, 49408, 768, 12, 12, lora_adapt=False, rank=-1,
                            design_details=design_details)
    sd = {k: v.detach().clone() for k, v in clip.state_dict().items()}
    # give the zero-initialised pieces some signal (class embedding etc. are already random; LayerNorm affine stays 1 / 0)
    g = torch.Generator().manual_seed(seed + 1)
    for k in list(sd):
        if k.endswith(("in_proj_bias", "in_proj.bias", "out_proj.bias", "c_fc.bias", "c_proj.bias")):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.02
    # build_model expects the OpenAI key names for the visual attention (it renames them itself, :866-876)
    return {k.replace("in_proj.weight", "in_proj_weight").replace("in_proj.bias", "in_proj_bias") if k.startswith("visual.") else k: v
            for k, v in sd.items()}


class Cascade:
    """The reference cascade, ready to run: ``sam`` (models/sam_maskdecoder_edge.py SAM with clip_model loaded)."""

    def __init__(self, sam, names, device):
        self.sam, self.class_names, self.device = sam, names, device

    @torch.no_grad()
    def run(self, inp: torch.Tensor, clip_image: torch.Tensor, clip_mask: torch.Tensor) -> Dict[str, torch.Tensor]:
        """demo.py:110-122: infer_test -> sigmoid -> alpha for the second CLIP pass -> predicted class."""
        sam = self.sam
        logits = sam.infer_test(inp.to(self.device), clip_image.to(self.device), clip_mask.to(self.device))
        pred_mask = torch.sigmoid(logits)
        alpha = F.interpolate(pred_mask, (336, 336), mode="bilinear", align_corners=False)
        _, _, pred_1, score = sam.clip_model(clip_image.to(self.device), alpha, train=False)
        return {"mask_logits": logits.float().cpu(), "class_id": pred_1.cpu(), "score": score.float().cpu()}


def build_cascade(device="cpu", seed: int = 0) -> Cascade:
    """SAM + CustomCLIP exactly as demo.py:78-88 builds them, random-init weights (seeded), eval mode."""
    root = cascade_root()
    if not root:
        raise FileNotFoundError("no reference cascade sources (neither /root/reference nor oracle/_ref/cascade)")
    import yaml
    with _reference_imports(root):
        sam_mod = importlib.import_module("models.sam_maskdecoder_edge")
        maple = importlib.import_module("cocotrainers.mapleAlphaCLIP")
        clip_model_mod = importlib.import_module("alpha_clip_rw.model")
        names = importlib.import_module("datasets.ovcamo_info.class_names")
        with open(os.path.join(root, "configs", "demo.yaml")) as f:
            config = yaml.load(f, Loader=yaml.FullLoader)
        cfg = DotDict(config["MAPLE_ALPHA_CLIP"])

        def load_clip_to_cpu(cfg_, device="cpu", alpha_vision_ckpt_pth=None):      # replaces the checkpoint download
            design = {"trainer": "MaPLe", "vision_depth": 0, "language_depth": 0, "vision_ctx": 0, "language_ctx": 0,
                      "maple_length": cfg_.TRAINER.MAPLE.N_CTX}
            sd = random_clip_state_dict(clip_model_mod, design, seed + 100)
            model = clip_model_mod.build_model(sd, lora_adapt=False, rank=-1, design_details=design).to(device)
            return model.float()

        maple.load_clip_to_cpu = load_clip_to_cpu
        torch.manual_seed(seed + 200)
        maple_clip_model = maple.TestMaPLeAlphaCLIP(cfg, names.TRAIN_CLASS_NAMES, names.TEST_CLASS_NAMES).model
        torch.manual_seed(seed + 300)
        args = config["model"]["args"]
        sam_cls = sam_mod.SAM
        # SAM.__init__ picks cuda when available (sam_maskdecoder_edge.py:116) and torch.load()s the shipped text features,
        # which were saved from a CUDA device, without map_location (:177-182): give it one on a CPU-only machine
        real_load = torch.load
        if not torch.cuda.is_available():
            torch.load = lambda f, *a, **k: real_load(f, *a, **{**k, "map_location": "cpu"})
        try:
            sam = sam_cls(**args)
        finally:
            torch.load = real_load
        sam.device = torch.device(device)
        sam.train_text_features = sam.train_text_features.to(sam.device)
        sam.test_text_features = sam.test_text_features.to(sam.device)
        sam.load_mapleAlphaCLIP(maple_clip_model)
    # seeded weights: the encoder / decoder with this repository's synthetic recipe, the small heads with torch's RNG
    from camouflaged_vlm_b200.config import PRESETS
    from camouflaged_vlm_b200.synth import seeded_state_dict
    from oracle import mask_decoder_oracle as mdo
    sam.image_encoder.load_state_dict(seeded_state_dict(PRESETS["vit_h"], seed), strict=True)
    dsd = mdo.seeded_decoder_state(seed)
    sam.mask_decoder.load_state_dict({k: v for k, v in dsd.items() if not k.startswith("probe.")}, strict=True)
    g = torch.Generator().manual_seed(seed + 400)
    with torch.no_grad():
        for m in list(sam.sam_visual_proj) + list(sam.sam_text_proj):
            if isinstance(m, nn.LayerNorm):
                m.weight.fill_(1.0)
                m.bias.zero_()
            else:
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * 0.05)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
        sam.no_mask_embed.weight.copy_(torch.randn(sam.no_mask_embed.weight.shape, generator=g) * 0.05)
        sam.pe_layer.positional_encoding_gaussian_matrix.copy_(torch.randn(2, 128, generator=g))
    sam = sam.to(device).eval()
    sam.clip_model.to(device).eval()
    return Cascade(sam, list(names.TEST_CLASS_NAMES), torch.device(device))


def synthetic_sample(seed: int = 0):
    """ImageNet-normalised 1024^2 input, CLIP-normalised 336^2 image, all-ones alpha after
    mask_transform = (1 - 0.5) / 0.26 (alpha_clip_rw/alpha_clip.py:88-94)."""
    from camouflaged_vlm_b200.config import PRESETS
    from camouflaged_vlm_b200.synth import _normal, seeded_images
    inp = seeded_images(PRESETS["vit_h"], 1, seed)
    clip_image = torch.from_numpy(_normal(f"clip_image.{seed}", (1, 3, 336, 336), seed))
    clip_mask = torch.full((1, 1, 336, 336), (1.0 - 0.5) / 0.26)
    return inp, clip_image, clip_mask


def swap_in_b200_modules(cascade: Cascade) -> None:
    """Replace the three classes with this repository's, in place, carrying the weights over by state_dict (strict):
    what a user of the reference does to switch."""
    from functools import partial

    from camouflaged_vlm_b200 import ImageEncoderViT
    from camouflaged_vlm_b200 import alpha_clip as ac
    from camouflaged_vlm_b200 import mask_decoder as md
    sam = cascade.sam
    dev = cascade.device
    ref_enc = sam.image_encoder
    enc = ImageEncoderViT(img_size=1024, patch_size=16, in_chans=3, embed_dim=1280, depth=32, num_heads=16, mlp_ratio=4,
                          out_chans=256, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), act_layer=nn.GELU,
                          use_rel_pos=True, rel_pos_zero_init=True, window_size=14, global_attn_indexes=[7, 15, 23, 31])
    enc.load_state_dict(ref_enc.state_dict(), strict=True)
    sam.image_encoder = enc.to(dev).eval()
    dec = md.build_mask_decoder()
    dec.load_state_dict(sam.mask_decoder.state_dict(), strict=True)
    sam.mask_decoder = dec.to(dev).eval()
    ref_vis = sam.clip_model.image_encoder
    n_ctx = int(ref_vis.transformer.resblocks[0].compound_prompt_nctx) if hasattr(ref_vis.transformer.resblocks[0], "compound_prompt_nctx") else 4
    tower = ac.VisionTransformer_MaPLe(336, 14, 1024, 24, 16, 768, design_details={"trainer": "MaPLe", "maple_length": n_ctx})
    tower.load_state_dict(ref_vis.state_dict(), strict=True)
    tower = tower.to(dev).eval()
    sam.clip_model.image_encoder = tower
    sam.clip_model.to(dev)
