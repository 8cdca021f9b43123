"""Pack the reference state_dict (image_encoder.py:65-130 parameter tree) into the tensors the
kernels consume (names consumed by csrc/encoder.cu; layouts in DESIGN.md "Packed weights").

All transformations are exact in real arithmetic:
  * K / V biases: ``b_k`` adds the same ``q.b_k`` to every logit of a softmax row and cancels;
    ``b_v`` passes through the convex combination unchanged, so ``proj(o + b_v) = proj(o) + W_p b_v``
    is folded into the proj bias.  K and V are therefore produced WITHOUT bias, which makes the
    zero padding of window_partition (image_encoder.py:518-528) equal to TMA/cp.async zero fill.
  * adapter: ``embedding_generator`` reads a scrambled view of (patch_embed), see
    ``oracle/encoder_oracle.py::prompts``; our stem keeps ``x0 = patch + pos_embed`` only, so the
    pos_embed part of that linear map is precomputed here (``emb.cst``).
  * prompt_{i+1} = shared_mlp(G_{i+1}) is folded into block i's lin2 (bias ``lin2.bf = b2 + b_s``).
  * norm1 (image_encoder.py:436) is folded around the GEMMs on either side of it:
    ``LN(x) W^T + b = rstd * (x W'^T - mu * c) + (W beta + b)`` with ``W' = W_qkv * gamma`` (column scale) and
    ``c[n] = sum_k W'[n, k]``.  ``c`` is summed over the ROUNDED operand the tensor cores see, so the
    mean term cancels exactly what the GEMM accumulates for a constant row.  (norm2 stays a kernel: its
    neighbours, proj and lin1, are the two GEMMs whose epilogues have no slack -- measured, DESIGN.md.)
``dtype=torch.float32`` yields the same tensors unrounded for the fp32 check path.
"""
from typing import Dict

import torch
import torch.nn.functional as F

from .config import EncoderConfig


def rank_pad(rank: int) -> int:
    return (rank + 7) // 8 * 8


def _resize_rel_pos(rel_pos: torch.Tensor, size: int) -> torch.Tensor:
    """get_rel_pos (image_encoder.py:567-579): linear interpolation when the table length differs."""
    want = 2 * size - 1
    if rel_pos.shape[0] == want:
        return rel_pos
    rp = F.interpolate(rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=want, mode="linear")
    return rp.reshape(-1, want).permute(1, 0)


def high_pass_matrices(n: int, rate: float, device, dtype=torch.bfloat16):
    """The low-pass projector of PromptGenerator.fft (image_encoder.py:332-353) for an n x n image:
    P[a, b] = (1/n) sum_{u in [-line, line)} exp(2 pi i u (a - b) / n),  line = int(sqrt(n*n*rate) // 2),
    so that  fft(x, rate) = |x - Re(P x P^T)| = |x - Pr x Pr^T + Pi x Pi^T|.
    Returns (pcat = [Pr; Pi] [2n, n], npr = -Pr, pi = Pi, the last two with their columns zero-padded to a
    multiple of 64: csrc/highpass.cu contracts them against a per-image padded block)."""
    line = int((n * n * rate) ** 0.5 // 2)
    u = torch.arange(-line, line, dtype=torch.float64)
    d = torch.arange(n, dtype=torch.float64)
    ang = 2.0 * torch.pi * u[:, None] * d[None, :] / n                 # [2*line, n]: first column of the circulant
    col_r, col_i = torch.cos(ang).sum(0) / n, torch.sin(ang).sum(0) / n
    idx = (torch.arange(n)[:, None] - torch.arange(n)[None, :]) % n    # P[a, b] = col[(a - b) mod n]
    pr, pi = col_r[idx], col_i[idx]
    to = lambda t: t.to(device=device, dtype=dtype).contiguous()
    pad = (n + 63) // 64 * 64 - n
    return to(torch.cat([pr, pi], 0)), to(F.pad(-pr, (0, pad))), to(F.pad(pi, (0, pad)))


@torch.no_grad()
def pack_weights(sd: Dict[str, torch.Tensor], cfg: EncoderConfig, device, dtype=torch.bfloat16
                 ) -> Dict[str, torch.Tensor]:
    f32 = torch.float32

    def g(name):
        return sd[name].detach().to(device=device, dtype=f32)

    def mat(t):      # GEMM operand
        return t.to(dtype).contiguous()

    def vec(t):      # epilogue / LayerNorm operand, always fp32
        return t.to(f32).contiguous()

    D, T, r, depth = cfg.embed_dim, cfg.tokens, cfg.rank, cfg.depth
    RP = rank_pad(r)
    kp = cfg.in_chans * cfg.patch_size ** 2
    hd = cfg.head_dim
    out: Dict[str, torch.Tensor] = {}

    out["patch.w"] = mat(g("patch_embed.proj.weight").reshape(D, kp))
    out["patch.b"] = vec(g("patch_embed.proj.bias"))
    pos = g("pos_embed").reshape(T, D) if cfg.use_abs_pos else None
    if pos is not None:
        out["pos"] = vec(pos)

    pg = "prompt_generator."
    out["hand.w"] = mat(g(pg + "prompt_generator.proj.weight").reshape(r, kp))
    hb = torch.zeros(RP, device=device, dtype=f32)
    hb[:r] = g(pg + "prompt_generator.proj.bias")
    out["hand.b"] = hb
    w_eg = g(pg + "embedding_generator.weight")                  # [r, D]
    wt = torch.zeros(D, RP, device=device, dtype=f32)
    wt[:, :r] = w_eg.t()
    out["emb.wt"] = wt
    cst = torch.zeros(T, RP, device=device, dtype=f32)
    cst[:, :r] = g(pg + "embedding_generator.bias")[None, :]
    if pos is not None:
        # scramble(pos)[t', c'] = pos_flat[c'*T + t']  ->  E[t', j] = sum_c' W[j, c'] * pos_flat[c'*T + t']
        scr = pos.reshape(-1).reshape(D, T).double()
        cst[:, :r] -= (w_eg.double() @ scr).t().to(f32)
    out["emb.cst"] = cst
    lw = torch.zeros(depth * RP, RP, device=device, dtype=f32)
    lb = torch.zeros(depth * RP, device=device, dtype=f32)
    for i in range(depth):
        lw[i * RP:i * RP + r, :r] = g(pg + f"lightweight_mlp_{i}.0.weight")
        lb[i * RP:i * RP + r] = g(pg + f"lightweight_mlp_{i}.0.bias")
    out["lw.w"] = mat(lw)
    out["lw.b"] = lb
    sw = torch.zeros(D, RP, device=device, dtype=f32)
    sw[:, :r] = g(pg + "shared_mlp.weight")
    out["shared.w"] = mat(sw)
    b_s = g(pg + "shared_mlp.bias")
    out["shared.b"] = vec(b_s)

    for i in range(depth):
        b, o = f"blocks.{i}.", f"blk.{i}."
        size = cfg.block_window(i) or cfg.grid      # Attention.input_size, image_encoder.py:422
        g1, be1 = g(b + "norm1.weight"), g(b + "norm1.bias")
        w_qkv = g(b + "attn.qkv.weight")
        out[o + "qkv.w"] = mat(w_qkv * g1[None, :])
        out[o + "qkv.c"] = vec(out[o + "qkv.w"].double().sum(1))
        qb = torch.zeros(3 * D, device=device, dtype=f32)
        b_v = torch.zeros(D, device=device, dtype=f32)
        if cfg.qkv_bias:
            full = g(b + "attn.qkv.bias")
            qb[:D] = full[:D]
            b_v = full[2 * D:]
        # W beta stays on K and V: it is present for real tokens only (the reference pads AFTER norm1)
        out[o + "qkv.b"] = vec(qb.double() + w_qkv.double() @ be1.double())
        if cfg.use_rel_pos:
            out[o + "relh"] = mat(_resize_rel_pos(g(b + "attn.rel_pos_h"), size))
            out[o + "relw"] = mat(_resize_rel_pos(g(b + "attn.rel_pos_w"), size))
        else:
            out[o + "relh"] = torch.zeros(2 * size - 1, hd, device=device, dtype=dtype)
            out[o + "relw"] = torch.zeros(2 * size - 1, hd, device=device, dtype=dtype)
        w_p = g(b + "attn.proj.weight")
        out[o + "proj.w"] = mat(w_p)
        out[o + "proj.b"] = vec((g(b + "attn.proj.bias").double() + w_p.double() @ b_v.double()).to(f32))
        out[o + "ln2.w"] = vec(g(b + "norm2.weight"))
        out[o + "ln2.b"] = vec(g(b + "norm2.bias"))
        out[o + "lin1.w"] = mat(g(b + "mlp.lin1.weight"))
        out[o + "lin1.b"] = vec(g(b + "mlp.lin1.bias"))
        out[o + "lin2.w"] = mat(g(b + "mlp.lin2.weight"))
        b2 = g(b + "mlp.lin2.bias")
        out[o + "lin2.b"] = vec(b2)
        out[o + "lin2.bf"] = vec(b2 + b_s)

    out["hp.pcat"], out["hp.npr"], out["hp.pi"] = high_pass_matrices(cfg.img_size, cfg.freq_nums, device, dtype)

    oc = cfg.out_chans
    out["neck.conv1.w"] = mat(g("neck.0.weight").reshape(oc, D))
    out["neck.ln1.w"] = vec(g("neck.1.weight"))
    out["neck.ln1.b"] = vec(g("neck.1.bias"))
    out["neck.conv2.w"] = mat(g("neck.2.weight").permute(0, 2, 3, 1).reshape(oc, 9 * oc))   # [n, (ky,kx), c]
    out["neck.ln2.w"] = vec(g("neck.3.weight"))
    out["neck.ln2.b"] = vec(g("neck.3.bias"))
    return out
