"""Adapter-only training path on the B200: kernel-level gradient checks against torch autograd in fp32,
and the gradients of every ``prompt_generator.*`` parameter through the nn.Module (the reference's training
configuration, train_ovcos_maskdecoder_edge.py:301-303) against goldens produced by the UNMODIFIED reference under
autograd (tests/golden/encoder_grads_*.npz, oracle/make_golden_grads.py).

Tolerance: the forward is bf16 (gate 1e-2 on the embedding); a gradient crosses the same 32 blocks a second time with
bf16 operands, gate rel-L2 <= 2e-2 per tensor group."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from camouflaged_vlm_b200 import PRESETS, build_encoder, ops           # noqa: E402
from camouflaged_vlm_b200.synth import seeded_images, seeded_state_dict  # noqa: E402
from oracle import encoder_oracle as orc                                 # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
GRAD_TOL = 2e-2


def _randn(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


def _rel(a, b):
    return orc.rel_l2(a.float().cpu(), b.float().cpu())


# ---------------------------------------------------------------------------------------------- attention backward
def attention_reference(qkv, relh, relw, batch, grid, window, heads, hd):
    """fp32 torch restatement of the kernel contract of cvlm_op_attention (bias-free K / V, zero padding AFTER the
    projection, padded keys counted in the softmax, rel-pos from the unscaled q: image_encoder.py:488-625)."""
    D = heads * hd
    S = window if window > 0 else grid
    x = qkv.float().reshape(batch, grid, grid, 3, heads, hd)
    pad = (S - grid % S) % S
    x = F.pad(x, (0, 0, 0, 0, 0, 0, 0, pad, 0, pad))
    gp = grid + pad
    nw = gp // S
    x = x.reshape(batch, nw, S, nw, S, 3, heads, hd).permute(5, 0, 1, 3, 6, 2, 4, 7).reshape(3, batch * nw * nw * heads, S * S, hd)
    q, k, v = x[0], x[1], x[2]
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    idx = torch.arange(S, device=qkv.device)[:, None] - torch.arange(S, device=qkv.device)[None, :] + (S - 1)
    rh, rw = relh.float()[idx], relw.float()[idx]                        # [S, S, hd]
    rq = q.reshape(-1, S, S, hd)
    rel_h = torch.einsum("bhwc,hkc->bhwk", rq, rh)
    rel_w = torch.einsum("bhwc,wkc->bhwk", rq, rw)
    attn = (attn.view(-1, S, S, S, S) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]).view(-1, S * S, S * S)
    o = attn.softmax(-1) @ v
    o = o.reshape(batch, nw, nw, heads, S, S, hd).permute(0, 1, 4, 2, 5, 3, 6).reshape(batch, gp, gp, D)
    return o[:, :grid, :grid].reshape(batch * grid * grid, D)


@pytest.mark.parametrize("batch,grid,window,heads,hd", [
    (2, 10, 4, 2, 32),       # tiny fixture geometry: 4x4 windows on a 10x10 grid (padded), head_dim 32
    (1, 18, 14, 2, 80),      # 14x14 windows on an 18x18 grid (padded to 28): the ViT-H window geometry, head_dim 80
    (1, 18, 14, 2, 64),
    (2, 16, 0, 2, 80),       # small global grid
    (1, 64, 14, 1, 80),      # full-size window layout: 25 windows, right / bottom edge padding
    (1, 64, 0, 2, 80),       # the 64x64 global block of ViT-H
    (1, 64, 0, 1, 64),
])
def test_attention_backward_matches_autograd(batch, grid, window, heads, hd):
    _attention_backward_case(batch, grid, window, heads, hd)


def test_attention_backward_generic_kernels_on_the_global_grid():
    """The 64x64 global blocks normally take the specialised kernels (attn_bwd_dq64 / dkv64); the generic pair must give
    the same gradients there ("attn_legacy" forces it, and the mma.sync forward with it)."""
    from camouflaged_vlm_b200 import _lib
    _lib.set_option("attn_legacy", 1)
    try:
        _attention_backward_case(1, 64, 0, 1, 80)
    finally:
        _lib.set_option("attn_legacy", 0)


@pytest.mark.parametrize("batch,grid,window,heads,hd", [(1, 64, 0, 2, 80), (2, 64, 14, 2, 80), (1, 64, 14, 1, 64),
                                                      (1, 28, 14, 2, 80), (1, 64, 0, 1, 64)])
def test_attention_backward_with_the_forward_log_sum_exp(batch, grid, window, heads, hd):
    """The training forward keeps every row's log-sum-exp (written by the tcgen05 attention kernels) so that the backward
    kernels skip their own sweep over S: same gradients as the self-contained backward, and the saved values equal the
    ones the backward computes for itself."""
    from camouflaged_vlm_b200 import _lib
    D = heads * hd
    S = window if window > 0 else grid
    rows = batch * grid * grid
    qkv = _randn(rows, 3 * D, seed=21, scale=1.0, dtype=torch.bfloat16)
    relh = _randn(2 * S - 1, hd, seed=22, scale=0.3, dtype=torch.bfloat16)
    relw = _randn(2 * S - 1, hd, seed=23, scale=0.3, dtype=torch.bfloat16)
    dout = _randn(rows, D, seed=24, dtype=torch.bfloat16)
    o, lse = ops.attention_lse(qkv, relh, relw, batch, grid, window, heads, hd)
    assert torch.equal(o, ops.attention(qkv, relh, relw, batch, grid, window, heads, hd))
    assert torch.isfinite(lse).all()
    # the backward's own log-sum-exp (global blocks write it to the scratch)
    dqkv0 = torch.zeros_like(qkv)
    lse0 = torch.full((rows, heads), float("nan"), device="cuda")
    dsum = torch.empty_like(lse0)
    _lib.check(_lib.load().cvlm_op_attention_backward(qkv.data_ptr(), o.data_ptr(), dout.data_ptr(), dqkv0.data_ptr(),
                                                      lse0.data_ptr(), dsum.data_ptr(), relh.data_ptr(), relw.data_ptr(),
                                                      batch, grid, window, heads, hd, _lib.stream_ptr()), "attention_backward")
    if window == 0:
        assert (lse - lse0).abs().max().item() < 2e-3          # log2 units; the forward's polynomial exponentials
    got = ops.attention_backward(qkv, o, dout, relh, relw, batch, grid, window, heads, hd, lse=lse)
    torch.cuda.synchronize()
    assert _rel(got, dqkv0) < 2e-3
    if window == 0:
        # global blocks: with the forward's per-query rel-pos terms the dK / dV kernel projects nothing itself
        o2, lse2, tw, th = ops.attention_lse(qkv, relh, relw, batch, grid, window, heads, hd, tables=True)
        assert torch.equal(o2, o) and torch.equal(lse2, lse)
        q = qkv[:, :D].float().reshape(batch, grid, grid, heads, hd)
        log2e = 1.4426950408889634
        idx = (torch.arange(grid, device="cuda")[:, None] - torch.arange(grid, device="cuda")[None, :] + grid - 1)
        want_tw = torch.einsum("byxhd,xkd->byxhk", q, relw.float()[idx]) * log2e          # [b, y, x, h, kx]
        want_th = torch.einsum("byxhd,ykd->byxhk", q, relh.float()[idx]) * log2e          # [b, y, x, h, ky]
        # (the T_w table carries the row's log-sum-exp: tw = T_w - lse, so that P = exp2(s * scale + tw + th))
        assert _rel(tw.reshape(batch, grid, grid, heads, 64) + lse.reshape(batch, grid, grid, heads, 1), want_tw) < 1e-5
        assert _rel(th.reshape(batch, grid, grid, heads, 64), want_th) < 1e-5
        # bwd_tc: 0 = mma.sync kernels with the tables, 1 = dK / dV on tcgen05, 2 = dQ too (attention_bwd_tc.cu; the default)
        for tc in (0, 1, 2):
            _lib.set_option("bwd_tc", tc)
            try:
                got_t = ops.attention_backward(qkv, o, dout, relh, relw, batch, grid, window, heads, hd, lse=lse, tables=(tw, th))
                torch.cuda.synchronize()
            finally:
                _lib.set_option("bwd_tc", 2)
            if tc < 2:
                assert torch.equal(got_t[:, :D], got[:, :D])         # dQ: same kernel
            assert _rel(got_t[:, :D], dqkv0[:, :D]) < 3e-3, f"dQ, bwd_tc={tc}"
            assert _rel(got_t[:, D:2 * D], dqkv0[:, D:2 * D]) < 3e-3, f"dK, bwd_tc={tc}"
            assert _rel(got_t[:, 2 * D:], dqkv0[:, 2 * D:]) < 3e-3, f"dV, bwd_tc={tc}"
        got = got_t
    leaf = qkv.float().requires_grad_(True)
    (attention_reference(leaf, relh, relw, batch, grid, window, heads, hd) * dout.float()).sum().backward()
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        err = _rel(got[:, sl], leaf.grad[:, sl])
        assert err < GRAD_TOL, f"{name}: rel-L2 {err:.3e}"


def test_attention_lse_refuses_the_fallback_kernel():
    qkv = _randn(4 * 16 * 16, 3 * 64, seed=31, dtype=torch.bfloat16)
    rel = _randn(2 * 16 - 1, 32, seed=32, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.attention_lse(qkv, rel, rel, 4, 16, 0, 2, 32)


def _attention_backward_case(batch, grid, window, heads, hd):
    D = heads * hd
    S = window if window > 0 else grid
    rows = batch * grid * grid
    qkv = _randn(rows, 3 * D, seed=1, scale=1.0, dtype=torch.bfloat16)
    relh = _randn(2 * S - 1, hd, seed=2, scale=0.3, dtype=torch.bfloat16)
    relw = _randn(2 * S - 1, hd, seed=3, scale=0.3, dtype=torch.bfloat16)
    dout = _randn(rows, D, seed=4, dtype=torch.bfloat16)
    o = ops.attention(qkv, relh, relw, batch, grid, window, heads, hd)
    leaf = qkv.float().requires_grad_(True)
    want_o = attention_reference(leaf, relh, relw, batch, grid, window, heads, hd)
    assert _rel(o, want_o) < 1e-2
    (want_o * dout.float()).sum().backward()
    got = ops.attention_backward(qkv, o, dout, relh, relw, batch, grid, window, heads, hd)
    torch.cuda.synchronize()
    want = leaf.grad
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        err = _rel(got[:, sl], want[:, sl])
        assert err < GRAD_TOL, f"{name}: rel-L2 {err:.3e}"


# ---------------------------------------------------------------------------------------------- row / wgrad kernels
@pytest.mark.parametrize("cols", [64, 256, 320, 1280])
def test_ln_bwd_all_input_kinds(cols):
    rows, eps = 517, 1e-6
    x = _randn(rows, cols, seed=10, scale=2.0) + 0.3
    gamma = 1.0 + _randn(cols, seed=11, scale=0.1)
    dy = _randn(rows, cols, seed=12)
    res = _randn(rows, cols, seed=13)
    leaf = x.clone().requires_grad_(True)
    y = F.layer_norm(leaf, (cols,), gamma, None, eps)
    (y * dy).sum().backward()
    want = leaf.grad + res
    mean, var = x.mean(1, keepdim=True), x.var(1, unbiased=False, keepdim=True)
    rstd = (var + eps).rsqrt()
    # kind 2: fp32 x, gamma applied in the kernel
    dx, dxb = ops.ln_bwd(2, dy, x, gamma=gamma, residual=res, eps=eps)
    assert _rel(dx, want) < 1e-5 and _rel(dxb, want) < 4e-3
    # kind 1: saved bf16 xhat + rstd, gamma folded by the caller
    xhat = ((x - mean) * rstd).to(torch.bfloat16)
    dx1, _ = ops.ln_bwd(1, (dy * gamma).contiguous(), xhat, stat=rstd.reshape(-1).contiguous(), residual=res)
    assert _rel(dx1, want) < 4e-3
    # kind 0: raw bf16 x + (rstd, -mean * rstd) rows (the norm1 fold)
    xb = x.to(torch.bfloat16)
    rows_stat = torch.cat([rstd, -mean * rstd], 1).contiguous()
    dx0, _ = ops.ln_bwd(0, (dy * gamma).contiguous(), xb, stat=rows_stat, residual=res)
    assert _rel(dx0, want) < 6e-3
    # in-place on the residual stream (dx aliases residual), as the schedule runs it
    inplace = res.clone()
    from camouflaged_vlm_b200 import _lib
    _lib.check(_lib.load().cvlm_op_ln_bwd(2, dy.data_ptr(), gamma.data_ptr(), x.data_ptr(), None, inplace.data_ptr(),
                                          inplace.data_ptr(), None, rows, cols, eps, _lib.stream_ptr()))
    assert torch.equal(inplace, dx)


@pytest.mark.parametrize("m,i_cols,j_cols,ones", [(100, 64, 8, True), (4096, 1280, 40, True), (5000, 768, 40, False),
                                                  (4100, 40, 40, True), (333, 32, 8, True), (2048 * 3 + 5, 128, 16, False)])
def test_wgrad_tn(m, i_cols, j_cols, ones):
    a = _randn(m, i_cols, seed=20, dtype=torch.bfloat16)
    big = _randn(m, 3 * j_cols, seed=21, dtype=torch.bfloat16)
    b = big[:, j_cols:2 * j_cols]                                      # a column slice: row stride != J
    got = ops.wgrad_tn(a, b, ones=ones)
    want = a.float().t() @ b.float()
    assert _rel(got[:, :j_cols], want) < 1e-5
    if ones:
        assert _rel(got[:, j_cols], a.float().sum(0)) < 1e-5
    assert torch.equal(got, ops.wgrad_tn(a, b, ones=ones))             # deterministic (one writer per partial)


@pytest.mark.parametrize("m,n,k", [(300, 5120, 1280), (4096, 40, 1280), (1000, 8, 64), (4096, 1280, 320)])
def test_gemm_dgelu_epilogue(m, n, k):
    a = _randn(m, k, seed=30, dtype=torch.bfloat16)
    b = _randn(n, k, seed=31, scale=0.05, dtype=torch.bfloat16)
    ldaux = 3 * n if n % 8 == 0 else n
    aux_full = _randn(m, ldaux, seed=32, scale=1.5, dtype=torch.bfloat16)
    aux = aux_full[:, n:2 * n] if ldaux != n else aux_full
    got = ops.gemm_bf16_dgelu(a, b, aux, n=n)
    want = (a.float() @ b.float().t()) * aux.float()
    assert _rel(got, want) < 4e-3


@pytest.mark.parametrize("m,n,k", [(300, 5120, 1280), (4096, 40, 1280), (1000, 64, 64)])
def test_gelu_layer_keeps_its_derivative_for_the_backward(m, n, k):
    """Training forward of a Linear -> GELU layer: the bf16 GEMM's second output is gelu'(pre-activation) (what the
    backward-of-GELU epilogue multiplies by), computed from the same erf / exp as the GELU itself."""
    a = _randn(m, k, seed=33, dtype=torch.bfloat16)
    b = _randn(n, k, seed=34, scale=0.05, dtype=torch.bfloat16)
    bias = _randn(n, seed=35, scale=0.5)
    keep = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    out = ops.gemm_bf16(a, b, bias=bias, out_dtype=torch.bfloat16, gelu=True, copy_bf16=keep)
    h = (a.float() @ b.float().t() + bias).requires_grad_(True)
    act = F.gelu(h)
    act.sum().backward()
    assert _rel(out, act.detach()) < 4e-3
    assert _rel(keep, h.grad) < 4e-3
    assert torch.equal(out, ops.gemm_bf16(a, b, bias=bias, out_dtype=torch.bfloat16, gelu=True))   # same GELU bits either way


# ---------------------------------------------------------------------------------------------- whole module
def _trainable_encoder(cfg, seed=0):
    m = build_encoder(cfg)
    m.load_state_dict(seeded_state_dict(cfg, seed), strict=True)
    m = m.cuda().train()
    for n, p in m.named_parameters():            # the reference's freeze rule, train_ovcos_maskdecoder_edge.py:301-303
        p.requires_grad_("prompt_generator" in n)
    return m


def _check_against_golden(name, batch, with_interm=True, tol=GRAD_TOL):
    cfg = PRESETS[name]
    gold = np.load(os.path.join(GOLDEN, f"encoder_grads_{name}.npz"))
    m = _trainable_encoder(cfg)
    x = seeded_images(cfg, batch, 0).cuda()
    c_out, c_int = orc.cotangents(cfg, batch, 0)
    out, interm = m(x, interm=True)
    assert out.requires_grad
    loss = (out * c_out.cuda()).sum()
    if with_interm:
        for t, c in zip(interm, c_int):
            loss = loss + (t * c.cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    # the loss is a sum of ~1e6 random-signed terms: the bf16 error of the embedding (rel-L2 <= 1e-2) moves it by about
    # 1e-2 * ||out|| * rms(c), so only a loose check makes sense here; the gradients below are the parity statement
    noise = 4e-2 * float(out.detach().norm()) + sum(4e-2 * 0.05 * float(t.detach().norm()) for t in interm)
    assert abs(float(loss.detach()) - float(gold["loss"])) <= noise
    worst, each = {}, {}
    for n, p in m.named_parameters():
        if not p.requires_grad:
            assert p.grad is None
            continue
        want = torch.from_numpy(gold[n])
        assert p.grad is not None and tuple(p.grad.shape) == tuple(want.shape), n
        group = n.split(".")[1] if "lightweight" not in n else "lightweight_mlp"
        if group == "prompt_generator":
            group = "prompt_generator.proj"
        each[n] = orc.rel_l2(p.grad.cpu(), want)
        worst[group] = max(worst.get(group, 0.0), each[n])
    print(f"{name}: adapter gradient rel-L2 by group: " + ", ".join(f"{k} {v:.2e}" for k, v in worst.items()))
    bad = {k: f"{v:.3e}" for k, v in each.items() if not v < tol}
    assert not bad, bad
    return m


@pytest.mark.parametrize("name,batch", [("tiny", 2), ("small_hd80", 1), ("small_hd64", 1)])
def test_adapter_gradients_match_reference_autograd(name, batch):
    # "tiny" has a rank-2 adapter on a 64-wide trunk: its lightweight-MLP gradients are 2x2 matrices, sums of 200 bf16
    # products each, and their relative error is a coin toss around 3e-2 (measured 2e-2 at batch 1, 5.6e-2 at batch 2,
    # with batch-2 == sum of the batch-1 passes to 2e-3: test_training_contract_of_the_module).  It gates at 1e-1 and
    # checks the plumbing (shapes, padding of the rank to 8, two images); the 2e-2 gate is held on the head_dim-80 / 64
    # fixtures and at full ViT-H size below.
    _check_against_golden(name, batch, tol=1e-1 if name == "tiny" else GRAD_TOL)


def test_adapter_gradients_vit_h_match_reference_autograd():
    """Full size: ViT-H at 1024^2, one image, all 70 adapter tensors against the reference's autograd."""
    m = _check_against_golden("vit_h", 1)
    assert m.last_backward_launch_count > 0


def test_training_contract_of_the_module():
    cfg = PRESETS["tiny"]
    m = _trainable_encoder(cfg)
    x = seeded_images(cfg, 2, 0).cuda()
    # 1. a real optimisation step changes the forward (packed weights are rebuilt from the updated parameters)
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=0.5)
    out0 = m(x)
    out0.square().mean().backward()
    g0 = {n: p.grad.clone() for n, p in m.named_parameters() if p.requires_grad}
    assert all(torch.isfinite(v).all() for v in g0.values()) and any(float(v.abs().max()) > 0 for v in g0.values())
    opt.step()
    opt.zero_grad()
    with torch.no_grad():
        out1 = m(x)
    assert orc.rel_l2(out1, out0.detach()) > 1e-4
    # 2. the inference path and the training path compute the same forward
    m2 = _trainable_encoder(cfg)
    with torch.no_grad():
        inf = m2(x)
        inf_i = m2(x, interm=True)[0]
    assert torch.equal(m2(x, interm=True)[0].detach(), inf_i)      # same schedule (no prompt fold at the global blocks)
    assert orc.rel_l2(m2(x).detach(), inf) < 1e-2
    # 3. gradients do not depend on the batch split (images are independent): batch 2 == sum of two batch-1 passes
    m3 = _trainable_encoder(cfg)
    c = orc.cotangents(cfg, 2, 0)[0].cuda()
    (m3(x) * c).sum().backward()
    both = {n: p.grad.clone() for n, p in m3.named_parameters() if p.requires_grad}
    m3.zero_grad()
    for k in range(2):
        (m3(x[k:k + 1]) * c[k:k + 1]).sum().backward()
    for n, p in m3.named_parameters():
        if p.requires_grad:
            assert orc.rel_l2(p.grad, both[n]) < 2e-3, n
    # 4. trunk parameters that require grad: training mode refuses with the fix in the message, eval() detaches
    m4 = build_encoder(cfg)
    m4.load_state_dict(seeded_state_dict(cfg, 0), strict=True)
    m4 = m4.cuda().train()
    with pytest.raises(RuntimeError, match="prompt_generator"):
        m4(x)
    m4.eval()
    with pytest.warns(UserWarning):
        out = m4(x)
    assert not out.requires_grad
    with pytest.raises(RuntimeError, match="input"):
        m(x.clone().requires_grad_(True))
