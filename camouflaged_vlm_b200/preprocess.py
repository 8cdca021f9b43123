"""The reference's image input pipeline on the GPU (datasets/wrappers.py:22-27, 61):
``transforms.Resize((S, S))`` on a PIL image -> ``ToTensor`` -> ``Normalize(ImageNet mean / std)``, bit-exact
(csrc/preprocess.cu).  uint8 HWC images of any size in, the encoder's fp32 ``[B, 3, S, S]`` input out; the H2D copy
of a raw image is 4x (same size) to 12x+ (larger photographs) smaller than that of the fp32 tensor it replaces."""
import ctypes as C
from typing import Sequence, Union

import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # datasets/wrappers.py:25-26
IMAGENET_STD = (0.229, 0.224, 0.225)


def preprocess_u8(images: Union[torch.Tensor, Sequence[torch.Tensor]], size: int = 1024,
                  mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD) -> torch.Tensor:
    """``images``: one uint8 CUDA tensor [H, W, 3] or a sequence of them (sizes may differ).  Returns fp32
    [B, 3, size, size] on the same device, launched on the current stream.  No CPU fallback."""
    if isinstance(images, torch.Tensor):
        images = [images]
    if not images:
        raise ValueError("preprocess_u8: no images")
    dev = images[0].device
    lib = _lib.load()
    out = torch.empty(len(images), 3, size, size, device=dev, dtype=torch.float32)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    need = 0
    for im in images:
        if not im.is_cuda:
            raise RuntimeError("preprocess_u8 runs on CUDA tensors only (no CPU fallback)")
        if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3 or not im.is_contiguous():
            raise ValueError("preprocess_u8 expects contiguous uint8 [H, W, 3] images")
        need = max(need, _lib.check(lib.cvlm_op_preprocess_u8_scratch_bytes(im.shape[0], im.shape[1], size),
                                    "cvlm_op_preprocess_u8_scratch_bytes"))
    # one scratch per image: the launches of consecutive images are independent and may overlap their tails
    scratch = torch.empty(len(images), (need + 255) // 256 * 256, dtype=torch.uint8, device=dev)
    stream = _lib.stream_ptr(dev)
    with torch.cuda.device(dev):                # the C ABI launches on the CUDA current device
        for i, im in enumerate(images):
            _lib.check(lib.cvlm_op_preprocess_u8(im.data_ptr(), im.shape[0], im.shape[1], out[i].data_ptr(), size, m, s,
                                                 scratch[i].data_ptr(), scratch.shape[1], stream), "cvlm_op_preprocess_u8")
    return out
