"""Drop-in replacement for the reference's Alpha-CLIP visual tower with MaPLe deep prompts:
``VisionTransformer_MaPLe`` of alpha_clip_rw/model.py:507-563 -- same constructor, same parameter tree / state_dict keys
(``conv1.weight``, ``conv1_alpha.weight``, ``class_embedding``, ``positional_embedding``, ``ln_pre.*``,
``transformer.resblocks.{i}.{attn.in_proj, attn.out_proj, ln_1, mlp.c_fc, mlp.c_proj, ln_2}.*``, ``ln_post.*``, ``proj``),
same ``forward(x, alpha, shared_ctx, compound_deeper_prompts)`` -> features ``[B, output_dim]``
(cocotrainers/mapleAlphaCLIP.py:272, 286).

The forward runs csrc/alpha_clip.cu through the C ABI (include/cvlm.h ``cvlm_vit_*``): fp32 residual stream, bf16 tensor-core
operands.  There is no CPU path and no PyTorch fallback; the sub-modules are parameter containers.  ``return_attn=True``
(the last block's attention map, unused by the cascade) is not implemented and raises.
"""
import ctypes as C
from collections import OrderedDict
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib


class _Attention(nn.Module):
    def __init__(self, dim: int, num_heads: int) -> None:
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.in_proj = nn.Linear(dim, dim * 3)
        self.out_proj = nn.Linear(dim, dim)


class _ResidualAttentionBlock(nn.Module):
    """Parameter container for CustomResidualAttentionBlock_MaPLe (alpha_clip_rw/model.py:285-309)."""

    def __init__(self, d_model: int, n_head: int) -> None:
        super().__init__()
        self.attn = _Attention(d_model, n_head)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", nn.Identity()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = nn.LayerNorm(d_model)


class _Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int) -> None:
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[_ResidualAttentionBlock(width, heads) for _ in range(layers)])


def pack_vit_weights(sd: Dict[str, torch.Tensor], layers: int, device) -> Dict[str, torch.Tensor]:
    """Device tensors registered with ``cvlm_vit_set_tensor``: every fp32 parameter under its reference name; ``patch.w``
    bf16 [width, 4 P^2 (padded to 8)] = conv1.weight and conv1_alpha.weight flattened side by side; ``<linear>.weight.bf16``
    for the four projections of every block; ``proj.t`` bf16 [output_dim, width]."""
    out: Dict[str, torch.Tensor] = {}
    f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
    for k, v in sd.items():
        out[k] = f32(v)
    w = out["conv1.weight"].shape[0]
    pw = torch.cat([out["conv1.weight"].reshape(w, -1), out["conv1_alpha.weight"].reshape(w, -1)], dim=1)
    kp = (pw.shape[1] + 7) // 8 * 8
    full = torch.zeros(w, kp, device=device, dtype=torch.float32)
    full[:, :pw.shape[1]] = pw
    out["patch.w"] = full.to(torch.bfloat16).contiguous()
    for i in range(layers):
        p = f"transformer.resblocks.{i}."
        for n in ("attn.in_proj", "attn.out_proj", "mlp.c_fc", "mlp.c_proj"):
            out[p + n + ".weight.bf16"] = out[p + n + ".weight"].to(torch.bfloat16).contiguous()
    out["proj.t"] = out["proj"].t().to(torch.bfloat16).contiguous()
    return out


class VisionTransformer_MaPLe(nn.Module):
    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int,
                 lora_adapt: bool = False, rank: int = 16, design_details: Optional[dict] = None) -> None:
        super().__init__()
        if lora_adapt:
            raise NotImplementedError("the B200 tower implements the cascade's configuration (lora_adapt=False)")
        if not design_details or "maple_length" not in design_details:
            raise ValueError("design_details['maple_length'] (number of MaPLe context tokens) is required")
        self.input_resolution = input_resolution
        self.output_dim = output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        self.conv1_alpha = nn.Conv2d(1, width, kernel_size=patch_size, stride=patch_size, bias=False)
        self.VPT_shallow = True
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.prompt_till_layer_visual = 0
        self.transformer = _Transformer(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self._n_ctx = int(design_details["maple_length"])
        self._patch, self._width, self._layers, self._heads = patch_size, width, layers, heads
        self.max_chunk = 64
        self._handle: Optional[int] = None
        self._packed = None
        self._packed_key = None
        self._workspace: Optional[torch.Tensor] = None
        self.last_launch_count = 0

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in ("_handle", "_packed", "_packed_key", "_workspace"):
            state[k] = None
        return state

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().cvlm_vit_destroy(self._handle)
        except Exception:
            pass

    def _weights_key(self, device):
        return (str(device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _get_workspace(self, need: int, dev) -> torch.Tensor:
        # a CUDA graph that captured the old buffer keeps its own reference (graphed.py), so replacing is safe
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != dev:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        return self._workspace

    def _ensure_ready(self, device) -> int:
        lib = _lib.load()
        if self._handle is None:
            cc = _lib.VitConfigC()
            cc.input_resolution, cc.patch_size, cc.width = self.input_resolution, self._patch, self._width
            cc.layers, cc.heads, cc.output_dim, cc.n_ctx = self._layers, self._heads, self.output_dim, self._n_ctx
            cc.ln_eps = float(self.ln_pre.eps)
            h = C.c_void_p()
            _lib.check(lib.cvlm_vit_create(C.byref(cc), C.byref(h)), "cvlm_vit_create")
            self._handle = h.value
        key = self._weights_key(device)
        if self._packed is None or self._packed_key != key:
            self._packed = pack_vit_weights(dict(self.state_dict()), self._layers, device)
            for name, t in self._packed.items():
                _lib.check(lib.cvlm_vit_set_tensor(self._handle, name.encode(), t.data_ptr(), t.numel() * t.element_size()),
                           f"vit set_tensor({name})")
            self._packed_key = key
        return self._handle

    def forward(self, x: torch.Tensor, alpha: torch.Tensor = None, shared_ctx: torch.Tensor = None,
                compound_deeper_prompts: Optional[List[torch.Tensor]] = None, return_attn: bool = False) -> torch.Tensor:
        if return_attn:
            raise NotImplementedError("return_attn=True is not implemented by the B200 tower")
        if not x.is_cuda:
            raise RuntimeError("VisionTransformer_MaPLe (B200) needs CUDA tensors: there is no CPU path")
        B, R = x.shape[0], self.input_resolution
        if tuple(x.shape[1:]) != (3, R, R) or alpha is None or tuple(alpha.shape) != (B, 1, R, R):
            raise ValueError(f"expected images [B,3,{R},{R}] and alpha [B,1,{R},{R}]")
        if shared_ctx is None or tuple(shared_ctx.shape) != (self._n_ctx, self._width):
            raise ValueError(f"shared_ctx must be [{self._n_ctx},{self._width}]")
        if torch.is_grad_enabled() and (x.requires_grad or alpha.requires_grad or shared_ctx.requires_grad or
                                        any(d.requires_grad for d in (compound_deeper_prompts or []))):
            # parameters that merely have requires_grad=True (the default of a freshly built module) do not block an
            # inference call; a differentiable INPUT does, because no gradient would reach it
            raise RuntimeError("VisionTransformer_MaPLe (B200) is inference only: an input requires grad; detach it or "
                               "call under torch.no_grad()")
        dev = x.device
        deep = list(compound_deeper_prompts or [])
        img = x.detach().float().contiguous()
        alp = alpha.detach().float().contiguous()
        ctx = shared_ctx.detach().float().contiguous()
        deep_t = torch.stack([d.detach().float() for d in deep]).contiguous() if deep else None
        if deep_t is not None and tuple(deep_t.shape[1:]) != (self._n_ctx, self._width):
            raise ValueError(f"every deep prompt must be [{self._n_ctx},{self._width}]")
        lib = _lib.load()
        h = self._ensure_ready(dev)
        out = torch.empty(B, self.output_dim, device=dev, dtype=torch.float32)
        launches = 0
        for s in range(0, B, self.max_chunk):
            e = min(B, s + self.max_chunk)
            need = _lib.check(lib.cvlm_vit_workspace_bytes(h, e - s), "cvlm_vit_workspace_bytes")
            ws = self._get_workspace(need, dev)
            with torch.cuda.device(dev):        # the C ABI launches on the CUDA current device
                _lib.check(lib.cvlm_vit_forward(h, img[s:e].data_ptr(), alp[s:e].data_ptr(), ctx.data_ptr(),
                                                deep_t.data_ptr() if deep_t is not None else None, len(deep), e - s,
                                                out[s:e].data_ptr(), ws.data_ptr(), ws.numel(),
                                                _lib.stream_ptr(dev)), "cvlm_vit_forward")
            launches += lib.cvlm_vit_last_launches(h)
        self.last_launch_count = launches
        return out.to(x.dtype) if x.dtype in (torch.float16, torch.bfloat16) else out
