"""Adapter-only training of the B200 encoder.

The reference fine-tunes only ``image_encoder.prompt_generator.*`` (plus its decoders) and freezes the rest of the
encoder (train_ovcos_maskdecoder_edge.py:301-303); ``loss.backward()`` then walks the autograd graph of
``self.image_encoder(input, interm=True)`` through all 32 blocks (models/sam_maskdecoder_edge.py:429-460).  Here that
graph is two C-ABI calls (include/cvlm.h): ``cvlm_encoder_forward_train`` keeps the activations the backward pass needs,
``cvlm_encoder_backward`` turns d(loss)/d(embedding) [and d(loss)/d(interm_k)] into the gradients of the adapter
parameters.  ``EncoderTrainFunction`` is the ``torch.autograd.Function`` that ``ImageEncoderViT.forward`` uses when grad
mode is on and only ``prompt_generator.*`` parameters require grad, so the reference training loop runs unmodified.

No trunk weight gradients and no input gradient exist on this path; ``ImageEncoderViT.forward`` raises if they are
asked for.
"""
import ctypes as C
from typing import Dict, List

import torch

from . import _lib
from .config import EncoderConfig
from .packing import rank_pad


@torch.no_grad()
def pack_backward_weights(sd: Dict[str, torch.Tensor], cfg: EncoderConfig, device, packed: Dict[str, torch.Tensor]
                          ) -> Dict[str, torch.Tensor]:
    """Transposed bf16 copies of the frozen weights for the dX GEMMs of the backward pass (names: include/cvlm.h).
    ``packed`` is the forward packing: the transposes are taken from the ROUNDED forward operands where one exists."""
    f32, bf = torch.float32, torch.bfloat16

    def g(name):
        return sd[name].detach().to(device=device, dtype=f32)

    D, r, depth, oc = cfg.embed_dim, cfg.rank, cfg.depth, cfg.out_chans
    RP = rank_pad(r)
    out: Dict[str, torch.Tensor] = {}
    for i in range(depth):
        b, o = f"blocks.{i}.", f"blk.{i}."
        out[o + "lin2.wt"] = packed[o + "lin2.w"].t().contiguous()                          # [4D, D]
        out[o + "lin1.wt"] = (g(b + "mlp.lin1.weight") * g(b + "norm2.weight")[None, :]).t().contiguous().to(bf)   # [D, 4D]
        out[o + "proj.wt"] = packed[o + "proj.w"].t().contiguous()                          # [D, D]
        out[o + "qkv.wt"] = packed[o + "qkv.w"].t().contiguous()                            # [D, 3D], norm1 gamma inside
    out["neck.conv1.wt"] = packed["neck.conv1.w"].t().contiguous()                          # [D, oc]
    # 3x3 conv backward = 3x3 conv of the output gradient with the flipped kernel, input / output channels swapped:
    # Wt[ci, (ky', kx'), co] = W[co, ci, 2 - ky', 2 - kx']
    w2 = g("neck.2.weight")
    out["neck.conv2.wt"] = w2.flip(2, 3).permute(1, 2, 3, 0).reshape(oc, 9 * oc).contiguous().to(bf)
    out["shared.wt"] = packed["shared.w"].t().contiguous()                                  # [RP, D]
    out["lw.wt"] = packed["lw.w"].t().contiguous()                                          # [RP, depth * RP]
    return out


def adapter_parameter_names(cfg: EncoderConfig) -> List[str]:
    pg = "prompt_generator."
    names = [pg + "shared_mlp.weight", pg + "shared_mlp.bias", pg + "embedding_generator.weight",
             pg + "embedding_generator.bias"]
    for i in range(cfg.depth):
        names += [pg + f"lightweight_mlp_{i}.0.weight", pg + f"lightweight_mlp_{i}.0.bias"]
    names += [pg + "prompt_generator.proj.weight", pg + "prompt_generator.proj.bias"]
    return names


def unpack_gradients(flat: torch.Tensor, offsets, jp: int, cfg: EncoderConfig) -> Dict[str, torch.Tensor]:
    """The library's gradient buffer (layout in include/cvlm.h) -> one tensor per adapter parameter, reference shapes."""
    D, r, depth = cfg.embed_dim, cfg.rank, cfg.depth
    RP = rank_pad(r)
    kp = cfg.in_chans * cfg.patch_size ** 2
    pg = "prompt_generator."
    shared = flat[offsets[0]:offsets[0] + D * jp].view(D, jp)
    lw = flat[offsets[1]:offsets[1] + depth * RP * jp].view(depth, RP, jp)
    hand = flat[offsets[2]:offsets[2] + kp * jp].view(kp, jp)
    dtsum = flat[offsets[3]:offsets[3] + RP * jp].view(RP, jp)
    eg = flat[offsets[4]:offsets[4] + D * RP].view(D, RP)
    out = {pg + "shared_mlp.weight": shared[:, :r].contiguous(), pg + "shared_mlp.bias": shared[:, RP].contiguous(),
           pg + "embedding_generator.weight": eg[:, :r].t().contiguous(),
           pg + "embedding_generator.bias": dtsum[:r, RP].contiguous(),
           pg + "prompt_generator.proj.weight": hand[:, :r].t().reshape(r, cfg.in_chans, cfg.patch_size, cfg.patch_size).contiguous(),
           pg + "prompt_generator.proj.bias": dtsum[:r, RP].clone()}
    for i in range(depth):
        out[pg + f"lightweight_mlp_{i}.0.weight"] = lw[i, :r, :r].contiguous()
        out[pg + f"lightweight_mlp_{i}.0.bias"] = lw[i, :r, RP].contiguous()
    return out


class EncoderTrainFunction(torch.autograd.Function):
    """forward(module, x, *adapter_params) -> (embedding NCHW, interm_0 NHWC, ...).  The adapter parameters are passed
    only so that autograd routes their gradients; the forward reads the module's packed copies."""

    @staticmethod
    def forward(ctx, module, x, *params):
        cfg = module.config
        dev = x.device
        lib = _lib.load()
        ctx.set_materialize_grads(False)       # outputs the loss does not use arrive as None, not as zero tensors
        module._ensure_ready(dev)
        module._ensure_backward_ready(dev)
        x = x.detach().to(torch.float32).contiguous()
        B, g, D = x.shape[0], cfg.grid, cfg.embed_dim
        out = torch.empty(B, cfg.out_chans, g, g, device=dev, dtype=torch.float32)
        n_glob = sum(1 for i in range(cfg.depth) if cfg.block_window(i) == 0)
        interms = [torch.empty(B, g, g, D, device=dev, dtype=torch.float32) for _ in range(n_glob)]
        chunks = []
        launches = 0
        stream = _lib.stream_ptr(dev)
        for s in range(0, B, module.max_chunk):
            e = min(B, s + module.max_chunk)
            ws = module._get_workspace(e - s, dev)
            need = _lib.check(lib.cvlm_encoder_saved_bytes(module._handle, e - s), "cvlm_encoder_saved_bytes")
            saved = torch.empty(need, dtype=torch.uint8, device=dev)
            arr = (C.c_void_p * n_glob)(*[t[s:e].data_ptr() for t in interms])
            with torch.cuda.device(dev):
                _lib.check(lib.cvlm_encoder_forward_train(module._handle, x[s:e].data_ptr(), e - s, out[s:e].data_ptr(),
                                                          C.cast(arr, C.POINTER(C.c_void_p)), saved.data_ptr(), saved.numel(),
                                                          ws.data_ptr(), ws.numel(), stream), "cvlm_encoder_forward_train")
            launches += lib.cvlm_encoder_last_launch_count(module._handle)
            chunks.append((s, e, saved))
        module.last_launch_count = launches
        ctx.module, ctx.chunks, ctx.n_glob = module, chunks, n_glob
        ctx.needs = [p is not None and p.requires_grad for p in params]
        return (out,) + tuple(interms)

    @staticmethod
    def backward(ctx, dout, *dinterms):
        module = ctx.module
        cfg = module.config
        lib = _lib.load()
        dev = dout.device if dout is not None else next(t for t in dinterms if t is not None).device
        g, D = cfg.grid, cfg.embed_dim
        nfloat = _lib.check(lib.cvlm_encoder_grad_floats(module._handle), "cvlm_encoder_grad_floats")
        offs = (C.c_longlong * 5)()
        jp = C.c_int()
        _lib.check(lib.cvlm_encoder_grad_layout(module._handle, offs, C.byref(jp)), "cvlm_encoder_grad_layout")
        total = None
        launches = 0
        stream = _lib.stream_ptr(dev)
        for s, e, saved in ctx.chunks:
            b = e - s
            if dout is None:
                d_nhwc = torch.zeros(b, g, g, cfg.out_chans, device=dev, dtype=torch.float32)
            else:
                d_nhwc = dout[s:e].detach().to(torch.float32).permute(0, 2, 3, 1).contiguous()
            dint = [None if t is None else t[s:e].detach().to(torch.float32).contiguous() for t in dinterms]
            arr = (C.c_void_p * max(1, ctx.n_glob))(*[None if t is None else t.data_ptr() for t in dint])
            need = _lib.check(lib.cvlm_encoder_backward_workspace_bytes(module._handle, b), "cvlm_encoder_backward_workspace_bytes")
            ws = module._get_backward_workspace(need, dev)
            flat = torch.empty(nfloat, device=dev, dtype=torch.float32)
            with torch.cuda.device(dev):
                _lib.check(lib.cvlm_encoder_backward(module._handle, b, d_nhwc.data_ptr(),
                                                     C.cast(arr, C.POINTER(C.c_void_p)) if ctx.n_glob else None,
                                                     saved.data_ptr(), saved.numel(), flat.data_ptr(), nfloat, ws.data_ptr(),
                                                     ws.numel(), stream), "cvlm_encoder_backward")
            launches += lib.cvlm_encoder_last_launch_count(module._handle)
            total = flat if total is None else total.add_(flat)
        module.last_backward_launch_count = launches
        ctx.chunks = None                      # release the saved activations
        grads = unpack_gradients(total, [int(v) for v in offs], jp.value, cfg)
        names = adapter_parameter_names(cfg)
        return (None, None) + tuple(grads[n] if need else None for n, need in zip(names, ctx.needs))
