"""Op-level Python wrappers over the C ABI (include/cvlm.h ``cvlm_op_*``): torch tensors in, torch
tensors out, launches on the current CUDA stream.  Used by the parity tests and for profiling single
kernels; the encoder itself goes through ``cvlm_encoder_forward``.  No fallbacks: CUDA tensors only."""
import functools
from typing import Optional

import torch

from . import _lib


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cvlm ops run on CUDA tensors only (no CPU fallback)")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _on_tensor_device(fn):
    """Run the wrapped op with its first tensor argument's device current: the C ABI launches on the CUDA current
    device and ``_lib.stream_ptr()`` then names torch's current stream of THAT device, so an op on cuda:1 tensors works
    whatever the caller's current device is.  All tensor arguments must live on one device."""
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        ts = [a for a in list(args) + list(kwargs.values()) if isinstance(a, torch.Tensor)]
        _need_cuda(*ts)
        if any(t.device != ts[0].device for t in ts):
            raise RuntimeError(f"{fn.__name__}: tensors live on different devices: {sorted({str(t.device) for t in ts})}")
        with torch.cuda.device(ts[0].device):
            return fn(*args, **kwargs)
    return wrapped


@_on_tensor_device
def gemm_bf16(a1: torch.Tensor, b1: torch.Tensor, a2: Optional[torch.Tensor] = None,
              b2: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
              residual: Optional[torch.Tensor] = None, res_row_mod: int = 0, out_dtype=torch.float32,
              gelu: bool = False, n: Optional[int] = None, out: Optional[torch.Tensor] = None,
              copy_bf16: Optional[torch.Tensor] = None, stat_out: Optional[torch.Tensor] = None,
              ln_rows: Optional[torch.Tensor] = None, ln_colsum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[m, n] = epilogue(a1 @ b1.T (+ a2 @ b2.T)).  a*: bf16 [m, k*] (last dim contiguous, any row stride),
    b*: bf16 [rows<=n, k*].  ``n`` defaults to b1.shape[0] rounded up to a multiple of 8.
    LayerNorm fold (cvlm_op_gemm_bf16_ln): ``copy_bf16`` [m, n] bf16 and ``stat_out`` [m, ceil(n/128), 2] fp32 are
    extra outputs of an fp32 GEMM; ``ln_rows`` [m, 2] (``ln_finalize`` of such statistics of the rows of a1) +
    ``ln_colsum`` [n] make a bf16 GEMM compute rstd * (acc - mean * colsum) + bias."""
    _need_cuda(a1, b1, a2, b2, bias, residual, out, copy_bf16, stat_out, ln_rows, ln_colsum)
    assert a1.dtype == torch.bfloat16 and b1.dtype == torch.bfloat16 and a1.stride(1) == 1 and b1.stride(1) == 1
    m, k1 = a1.shape
    b_rows = b1.shape[0]
    if n is None:
        n = (b_rows + 7) // 8 * 8
    if out is None:
        out = torch.empty(m, n, device=a1.device, dtype=out_dtype)
    k2 = 0
    if a2 is not None:
        assert a2.dtype == torch.bfloat16 and b2.dtype == torch.bfloat16 and a2.stride(1) == 1 and b2.stride(1) == 1
        k2 = a2.shape[1]
    lib = _lib.load()
    common = (a1.data_ptr(), a1.stride(0), b1.data_ptr(), b1.stride(0), k1,
              _ptr(a2), a2.stride(0) if a2 is not None else 0, _ptr(b2), b2.stride(0) if b2 is not None else 0, k2,
              m, n, b_rows, _ptr(bias), _ptr(residual), residual.stride(0) if residual is not None else 0, res_row_mod,
              out.data_ptr(), out.stride(0), 1 if out.dtype == torch.bfloat16 else 0, 1 if gelu else 0)
    if copy_bf16 is None and stat_out is None and ln_rows is None:
        _lib.check(lib.cvlm_op_gemm_bf16(*common, _lib.stream_ptr()), "cvlm_op_gemm_bf16")
        return out
    if copy_bf16 is not None:
        assert copy_bf16.dtype == torch.bfloat16 and copy_bf16.stride(1) == 1
    if stat_out is not None:
        assert stat_out.dtype == torch.float32 and stat_out.is_contiguous() and stat_out.shape == (m, -(-n // _lib.STAT_GROUP_COLS), 2)
    if ln_rows is not None:
        assert ln_rows.dtype == torch.float32 and ln_rows.is_contiguous() and ln_rows.shape == (m, 2)
        assert ln_colsum is not None and ln_colsum.dtype == torch.float32 and ln_colsum.numel() >= n
    _lib.check(lib.cvlm_op_gemm_bf16_ln(*common, _ptr(copy_bf16), copy_bf16.stride(0) if copy_bf16 is not None else 0,
                                        _ptr(stat_out), _ptr(ln_rows), _ptr(ln_colsum), _lib.stream_ptr()),
               "cvlm_op_gemm_bf16_ln")
    return out


@_on_tensor_device
def ln_finalize(stat: torch.Tensor, width: int, eps: float) -> torch.Tensor:
    """[rows, groups, 2] partial (sum, sum of squares) -> [rows, 2] = (rstd, -mean * rstd)."""
    _need_cuda(stat)
    assert stat.dtype == torch.float32 and stat.is_contiguous() and stat.dim() == 3 and stat.shape[2] == 2
    out = torch.empty(stat.shape[0], 2, device=stat.device, dtype=torch.float32)
    _lib.check(_lib.load().cvlm_op_ln_finalize(stat.data_ptr(), stat.shape[1], width, eps, out.data_ptr(), stat.shape[0],
                                               _lib.stream_ptr()), "cvlm_op_ln_finalize")
    return out


@_on_tensor_device
def attention(qkv: torch.Tensor, rel_pos_h: torch.Tensor, rel_pos_w: torch.Tensor, batch: int, grid: int,
              window: int, heads: int, head_dim: int) -> torch.Tensor:
    _need_cuda(qkv, rel_pos_h, rel_pos_w)
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous()
    out = torch.empty(qkv.shape[0], heads * head_dim, device=qkv.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().cvlm_op_attention(qkv.data_ptr(), out.data_ptr(), rel_pos_h.contiguous().data_ptr(),
                                             rel_pos_w.contiguous().data_ptr(), batch, grid, window, heads, head_dim,
                                             _lib.stream_ptr()), "cvlm_op_attention")
    return out


@_on_tensor_device
def attention_lse(qkv: torch.Tensor, rel_pos_h: torch.Tensor, rel_pos_w: torch.Tensor, batch: int, grid: int,
                  window: int, heads: int, head_dim: int, tables: bool = False):
    """``attention`` that also returns every token's row log-sum-exp (fp32 [rows, heads], log2 domain): what the training
    forward keeps for the backward kernels.  Only for the shapes of the tcgen05 kernels (64 x 64 global, 14 x 14 windows)."""
    _need_cuda(qkv, rel_pos_h, rel_pos_w)
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous()
    out = torch.empty(qkv.shape[0], heads * head_dim, device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty(qkv.shape[0], heads, device=qkv.device, dtype=torch.float32)
    tw = th = None
    if tables:      # 64 x 64 global blocks: every query's rel-pos terms for the dK / dV kernel, [rows, heads, 64] each
        tw = torch.empty(qkv.shape[0], heads, 64, device=qkv.device, dtype=torch.float32)
        th = torch.empty_like(tw)
    _lib.check(_lib.load().cvlm_op_attention_lse(qkv.data_ptr(), out.data_ptr(), rel_pos_h.contiguous().data_ptr(),
                                                 rel_pos_w.contiguous().data_ptr(), batch, grid, window, heads, head_dim,
                                                 lse.data_ptr(), _ptr(tw), _ptr(th), _lib.stream_ptr()), "cvlm_op_attention_lse")
    return (out, lse, tw, th) if tables else (out, lse)


@_on_tensor_device
def attention_plain(qkv: torch.Tensor, batch: int, tokens: int, heads: int, head_dim: int) -> torch.Tensor:
    """softmax(q k^T / sqrt(head_dim)) v per image, no relative-position terms; qkv bf16 [batch * tokens, 3 * heads * hd]."""
    _need_cuda(qkv)
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape == (batch * tokens, 3 * heads * head_dim)
    out = torch.empty(qkv.shape[0], heads * head_dim, device=qkv.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().cvlm_op_attention_plain(qkv.data_ptr(), out.data_ptr(), batch, tokens, heads, head_dim,
                                                   _lib.stream_ptr()), "cvlm_op_attention_plain")
    return out


@_on_tensor_device
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, out_dtype=torch.bfloat16):
    _need_cuda(x, gamma, beta)
    assert x.dtype == torch.float32 and x.is_contiguous()
    rows, cols = x.shape
    out = torch.empty(rows, cols, device=x.device, dtype=out_dtype)
    _lib.check(_lib.load().cvlm_op_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(),
                                             1 if out_dtype == torch.bfloat16 else 0, rows, cols, eps,
                                             _lib.stream_ptr()), "cvlm_op_layernorm")
    return out


@_on_tensor_device
def layernorm2d_nchw(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, batch: int, tokens: int, eps: float):
    _need_cuda(x, gamma, beta)
    assert x.dtype == torch.float32 and x.is_contiguous()
    chans = x.shape[1]
    out = torch.empty(batch, chans, tokens, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().cvlm_op_layernorm2d_nchw(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(),
                                                    batch, tokens, chans, eps, _lib.stream_ptr()),
               "cvlm_op_layernorm2d_nchw")
    return out


@_on_tensor_device
def im2col_patch16(img: torch.Tensor) -> torch.Tensor:
    _need_cuda(img)
    assert img.dtype == torch.float32 and img.is_contiguous()
    b, c, h, _ = img.shape
    g = h // 16
    out = torch.empty(b * g * g, c * 256, device=img.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().cvlm_op_im2col_patch16(img.data_ptr(), out.data_ptr(), b, c, h, _lib.stream_ptr()),
               "cvlm_op_im2col_patch16")
    return out


@_on_tensor_device
def im2col_3x3(x: torch.Tensor, batch: int, grid: int) -> torch.Tensor:
    _need_cuda(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    c = x.shape[-1]
    out = torch.empty(batch * grid * grid, 9 * c, device=x.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().cvlm_op_im2col_3x3(x.data_ptr(), out.data_ptr(), batch, grid, c, _lib.stream_ptr()),
               "cvlm_op_im2col_3x3")
    return out


@_on_tensor_device
def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    out = torch.empty_like(x, dtype=torch.bfloat16)
    _lib.check(_lib.load().cvlm_op_cast_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _lib.stream_ptr()),
               "cvlm_op_cast_bf16")
    return out


@_on_tensor_device
def high_pass(images: torch.Tensor, pcat: torch.Tensor, npr: torch.Tensor, pi: torch.Tensor) -> torch.Tensor:
    """PromptGenerator.fft as two GEMMs with the low-pass projector (packing.high_pass_matrices)."""
    _need_cuda(images, pcat, npr, pi)
    assert images.dtype == torch.float32 and images.is_contiguous() and images.dim() == 4
    b, c, h, w = images.shape
    hp_ = (h + 63) // 64 * 64
    assert h == w and pcat.shape == (2 * h, h) and npr.shape == (h, hp_) and pi.shape == (h, hp_)
    lib = _lib.load()
    need = lib.cvlm_op_high_pass_scratch_bytes(b, c, h)
    _lib.check(need, "cvlm_op_high_pass_scratch_bytes")
    scratch = torch.empty(need, dtype=torch.uint8, device=images.device)
    out = torch.empty_like(images)
    _lib.check(lib.cvlm_op_high_pass(images.data_ptr(), out.data_ptr(), b, c, h, pcat.contiguous().data_ptr(),
                                     npr.contiguous().data_ptr(), pi.contiguous().data_ptr(), scratch.data_ptr(), need,
                                     _lib.stream_ptr()), "cvlm_op_high_pass")
    return out


@_on_tensor_device
def adapter_embed(x0, hand, w_t, cst, batch, tokens, dim, rank, rank_pad) -> torch.Tensor:
    _need_cuda(x0, hand, w_t, cst)
    out = torch.empty(batch * tokens, rank_pad, device=x0.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().cvlm_op_adapter_embed(x0.data_ptr(), hand.data_ptr(), w_t.data_ptr(), cst.data_ptr(),
                                                 out.data_ptr(), batch, tokens, dim, rank, rank_pad,
                                                 _lib.stream_ptr()), "cvlm_op_adapter_embed")
    return out


# ------------------------------------------------------------------ backward-pass kernels (adapter-only training)
@_on_tensor_device
def attention_backward(qkv: torch.Tensor, o: torch.Tensor, dout: torch.Tensor, rel_pos_h: torch.Tensor,
                       rel_pos_w: torch.Tensor, batch: int, grid: int, window: int, heads: int, head_dim: int,
                       lse: Optional[torch.Tensor] = None, tables=None) -> torch.Tensor:
    """d(loss)/d(qkv) (bf16 [batch*grid^2, 3*D]) of ``attention`` given ``o`` (its output) and ``dout``.  ``lse`` (from
    ``attention_lse``) lets the kernels skip their own log-sum-exp sweep."""
    for t in (qkv, o, dout, rel_pos_h, rel_pos_w):
        assert t.dtype == torch.bfloat16 and t.is_contiguous()
    rows = batch * grid * grid
    dqkv = torch.zeros_like(qkv)
    dsum = torch.empty(rows, heads, device=qkv.device, dtype=torch.float32)
    if lse is not None:
        assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.shape == (rows, heads)
        _lib.check(_lib.load().cvlm_op_attention_backward_lse(qkv.data_ptr(), o.data_ptr(), dout.data_ptr(), dqkv.data_ptr(),
                                                              lse.data_ptr(), dsum.data_ptr(), rel_pos_h.data_ptr(),
                                                              rel_pos_w.data_ptr(), batch, grid, window, heads, head_dim,
                                                              _ptr(tables[0]) if tables else None,
                                                              _ptr(tables[1]) if tables else None,
                                                              _lib.stream_ptr()), "cvlm_op_attention_backward_lse")
        return dqkv
    lse = torch.empty_like(dsum)
    _lib.check(_lib.load().cvlm_op_attention_backward(qkv.data_ptr(), o.data_ptr(), dout.data_ptr(), dqkv.data_ptr(),
                                                      lse.data_ptr(), dsum.data_ptr(), rel_pos_h.data_ptr(), rel_pos_w.data_ptr(),
                                                      batch, grid, window, heads, head_dim, _lib.stream_ptr()),
               "cvlm_op_attention_backward")
    return dqkv


@_on_tensor_device
def ln_bwd(kind: int, g: torch.Tensor, xsrc: torch.Tensor, stat: Optional[torch.Tensor] = None,
           gamma: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, eps: float = 1e-6):
    """LayerNorm backward (csrc/backward.cu).  Returns (dx fp32, bf16 copy)."""
    rows, cols = g.shape
    assert g.dtype == torch.float32 and g.is_contiguous() and xsrc.is_contiguous()
    dx = torch.empty_like(g)
    dxb = torch.empty(rows, cols, device=g.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().cvlm_op_ln_bwd(kind, g.data_ptr(), _ptr(gamma), xsrc.data_ptr(), _ptr(stat), _ptr(residual),
                                          dx.data_ptr(), dxb.data_ptr(), rows, cols, eps, _lib.stream_ptr()), "cvlm_op_ln_bwd")
    return dx, dxb


@_on_tensor_device
def wgrad_tn(a: torch.Tensor, b: torch.Tensor, ones: bool = False) -> torch.Tensor:
    """out[i, j] = sum_m a[m, i] * b[m, j]  (fp32 [I, jp]; column J holds sum_m a[m, i] when ``ones``)."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.stride(1) == 1 and b.stride(1) == 1
    m, i_cols = a.shape
    j_cols = b.shape[1]
    lib = _lib.load()
    jp, splits = lib.cvlm_op_wgrad_tn_jp(j_cols, int(ones)), lib.cvlm_op_wgrad_tn_splits(m)
    scratch = torch.empty(splits, i_cols, jp, device=a.device, dtype=torch.float32)
    out = torch.empty(i_cols, jp, device=a.device, dtype=torch.float32)
    _lib.check(lib.cvlm_op_wgrad_tn(a.data_ptr(), a.stride(0), i_cols, b.data_ptr(), b.stride(0), j_cols, int(ones), m,
                                    scratch.data_ptr(), out.data_ptr(), _lib.stream_ptr()), "cvlm_op_wgrad_tn")
    return out


@_on_tensor_device
def gemm_bf16_dgelu(a: torch.Tensor, b: torch.Tensor, aux: torch.Tensor, n: Optional[int] = None) -> torch.Tensor:
    """(a @ b.T) * aux, bf16 out: the backward of a Linear -> GELU pair into the GELU's input, with ``aux`` = gelu'(h) as
    the training forward keeps it (``gemm_bf16(..., out_dtype=bf16, gelu=True, copy_bf16=aux)``)."""
    assert a.dtype == b.dtype == aux.dtype == torch.bfloat16 and a.stride(1) == 1 and b.stride(1) == 1 and aux.stride(1) == 1
    m, k = a.shape
    n = n or (b.shape[0] + 7) // 8 * 8
    out = torch.empty(m, n, device=a.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().cvlm_op_gemm_bf16_dgelu(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), k, m, n, b.shape[0],
                                                   aux.data_ptr(), aux.stride(0), out.data_ptr(), out.stride(0),
                                                   _lib.stream_ptr()), "cvlm_op_gemm_bf16_dgelu")
    return out
