"""fp32 check path: the launch schedule of csrc/encoder.cu re-enacted with torch ops.

This is NOT the product path and is never selected automatically.  It exists to prove, in fp32 and
on any device, that the restructuring the CUDA path relies on is exact: un-partitioned windows with
zero K/V padding, bias folding, the scrambled embedding read with the pre-folded pos_embed term,
prompt_{i+1} folded into lin2, the 3x3 conv as a 9-tap GEMM.  north_star tolerance: embedding
relative L2 <= 1e-5 against the reference in fp32.

Every function takes the PACKED tensors of packing.pack_weights(dtype=float32) and the same
arguments as the C op of the same name (include/cvlm.h).
"""
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .config import EncoderConfig
from .packing import rank_pad


def high_pass(x: torch.Tensor, rate: float) -> torch.Tensor:
    """PromptGenerator.fft (image_encoder.py:332-353) via torch.fft: remove the centred low-frequency
    box of half-width int(sqrt(H*W*rate) // 2), return |real part|."""
    h, w = x.shape[-2:]
    line = int((h * w * rate) ** 0.5 // 2)
    spec = torch.fft.fft2(x.float(), norm="forward")
    fy = torch.fft.fftfreq(h, d=1.0 / h, device=x.device).round().long()   # signed integer frequencies
    fx = torch.fft.fftfreq(w, d=1.0 / w, device=x.device).round().long()
    low_y = (fy >= -line) & (fy < line)
    low_x = (fx >= -line) & (fx < line)
    keep = ~(low_y[:, None] & low_x[None, :])
    return torch.fft.ifft2(spec * keep, norm="forward").real.abs()


def gemm(a1, b1, a2=None, b2=None, bias=None, residual=None, res_row_mod=0, gelu=False):
    y = a1.float() @ b1.float().t()
    if a2 is not None:
        y = y + a2.float() @ b2.float().t()
    if bias is not None:
        y = y + bias
    if gelu:
        y = F.gelu(y)
    if residual is not None:
        if res_row_mod > 0:
            idx = torch.arange(y.shape[0], device=y.device) % res_row_mod
            y = y + residual[idx]
        else:
            y = y + residual
    return y


def im2col_patch16(img: torch.Tensor) -> torch.Tensor:
    b, c, h, _ = img.shape
    g = h // 16
    p = img.reshape(b, c, g, 16, g, 16).permute(0, 2, 4, 1, 3, 5)
    return p.reshape(b * g * g, c * 256)


def im2col_3x3(x: torch.Tensor, batch: int, g: int) -> torch.Tensor:
    c = x.shape[-1]
    xp = F.pad(x.reshape(batch, g, g, c), (0, 0, 1, 1, 1, 1))
    taps = [xp[:, ky:ky + g, kx:kx + g, :] for ky in range(3) for kx in range(3)]
    return torch.stack(taps, dim=3).reshape(batch * g * g, 9 * c)


def adapter_embed(x0, hand, w_t, cst, batch, tokens, dim, rank, rp):
    scr = x0.reshape(batch, dim, tokens).permute(0, 2, 1)            # flat[c'*T + t'] -> [t', c']
    t = scr @ w_t + cst[None] + hand.reshape(batch, tokens, rp)
    t[..., rank:] = 0
    return t.reshape(batch * tokens, rp)


def attention(qkv, relh, relw, batch, g, window, heads, hd):
    """Same contract as cvlm_op_attention: token-major qkv with bias-free K/V, zero padding columns."""
    d = heads * hd
    s = window if window > 0 else g
    nw = -(-g // s)
    gp = nw * s
    q, k, v = (qkv[:, i * d:(i + 1) * d].reshape(batch, g, g, heads, hd) for i in range(3))

    def windows(t):   # [B, g, g, H, hd] -> [B*nw*nw, H, s*s, hd] with zero padding
        t = F.pad(t, (0, 0, 0, 0, 0, gp - g, 0, gp - g))
        t = t.reshape(batch, nw, s, nw, s, heads, hd).permute(0, 1, 3, 5, 2, 4, 6)
        return t.reshape(batch * nw * nw, heads, s * s, hd)

    qw, kw, vw = windows(q.float()), windows(k.float()), windows(v.float())
    logits = (qw * hd ** -0.5) @ kw.transpose(-2, -1)
    idx = torch.arange(s, device=qkv.device)
    rel = idx[:, None] - idx[None, :] + (s - 1)                                  # [q, k] -> table row
    rq = qw.reshape(-1, heads, s, s, hd)
    rel_h = torch.einsum("bnhwc,hkc->bnhwk", rq, relh.float()[rel])
    rel_w = torch.einsum("bnhwc,wkc->bnhwk", rq, relw.float()[rel])
    logits = (logits.reshape(-1, heads, s, s, s, s) + rel_h[..., :, None] + rel_w[..., None, :]
              ).reshape(-1, heads, s * s, s * s)
    o = logits.softmax(-1) @ vw
    o = o.reshape(batch, nw, nw, heads, s, s, hd).permute(0, 1, 4, 2, 5, 3, 6).reshape(batch, gp, gp, d)
    return o[:, :g, :g, :].reshape(batch * g * g, d)


def gemm_ln(x, w_scaled, colsum, bias, eps, gelu=False):
    """LayerNorm folded around a GEMM the way the CUDA epilogue does it (packing.py):
    rstd * (x W'^T - mu * colsum) + bias  ==  LN(x; gamma, beta) W^T + b."""
    mu = x.mean(-1, keepdim=True)
    var = (x * x).mean(-1, keepdim=True) - mu * mu
    rstd = torch.rsqrt(var.clamp_min(0) + eps)
    y = rstd * (x.float() @ w_scaled.float().t() - mu * colsum[None, :]) + bias[None, :]
    return F.gelu(y) if gelu else y


def layernorm2d_nchw(x, w, b, batch, tokens, eps):
    y = F.layer_norm(x, (x.shape[-1],), w, b, eps)
    return y.reshape(batch, tokens, -1).permute(0, 2, 1).contiguous()


@torch.no_grad()
def encoder_forward_check(p: Dict[str, torch.Tensor], cfg: EncoderConfig, images: torch.Tensor,
                          interm: bool = False, ln_eps: float = 1e-6, highpass: Optional[torch.Tensor] = None,
                          stop_after_block: Optional[int] = None):
    batch = images.shape[0]
    g, T, D, depth = cfg.grid, cfg.tokens, cfg.embed_dim, cfg.depth
    rp = rank_pad(cfg.rank)
    M = batch * T
    hp = high_pass(images, cfg.freq_nums) if highpass is None else highpass
    x = gemm(im2col_patch16(images.float()), p["patch.w"], bias=p["patch.b"], residual=p.get("pos"), res_row_mod=T)
    hw = torch.zeros(rp, p["hand.w"].shape[1], device=x.device)
    hw[:cfg.rank] = p["hand.w"].float()
    hand = gemm(im2col_patch16(hp), hw, bias=p["hand.b"])
    t = adapter_embed(x, hand, p["emb.wt"], p["emb.cst"], batch, T, D, cfg.rank, rp)
    gall = gemm(t, p["lw.w"], bias=p["lw.b"], gelu=True)

    def gslice(i):
        return gall[:, i * rp:(i + 1) * rp]

    x = gemm(gslice(0), p["shared.w"], bias=p["shared.b"], residual=x)
    if stop_after_block is not None and stop_after_block < 0:
        return x
    interms: List[torch.Tensor] = []
    for i in range(depth):
        o = f"blk.{i}."
        qkv = gemm_ln(x, p[o + "qkv.w"], p[o + "qkv.c"], p[o + "qkv.b"], ln_eps)
        a = attention(qkv, p[o + "relh"], p[o + "relw"], batch, g, cfg.block_window(i), cfg.num_heads, cfg.head_dim)
        x = gemm(a, p[o + "proj.w"], bias=p[o + "proj.b"], residual=x)
        a = F.layer_norm(x, (D,), p[o + "ln2.w"], p[o + "ln2.b"], ln_eps)
        h = gemm(a, p[o + "lin1.w"], bias=p[o + "lin1.b"], gelu=True)
        want_interm = interm and cfg.block_window(i) == 0
        stop_here = stop_after_block == i
        last = i + 1 == depth
        if not last and not want_interm and not stop_here:
            x = gemm(h, p[o + "lin2.w"], gslice(i + 1), p["shared.w"], bias=p[o + "lin2.bf"], residual=x)
        else:
            x = gemm(h, p[o + "lin2.w"], bias=p[o + "lin2.b"], residual=x)
            if want_interm:
                interms.append(x.reshape(batch, g, g, D).permute(0, 3, 1, 2))
            if stop_here:
                return x
            if not last:
                x = gemm(gslice(i + 1), p["shared.w"], bias=p["shared.b"], residual=x)
    n1 = gemm(x, p["neck.conv1.w"])
    n1n = F.layer_norm(n1, (cfg.out_chans,), p["neck.ln1.w"], p["neck.ln1.b"], 1e-6)
    n2 = gemm(im2col_3x3(n1n, batch, g), p["neck.conv2.w"])
    out = layernorm2d_nchw(n2, p["neck.ln2.w"], p["neck.ln2.b"], batch, T, 1e-6).reshape(batch, cfg.out_chans, g, g)
    return (out, interms) if interm else out
