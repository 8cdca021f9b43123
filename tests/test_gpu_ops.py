"""Kernel-level parity on the B200: every cvlm_op_* (through the C ABI) against a torch fp32
restatement of the same op on the same (bf16-rounded) inputs.  Tolerances are stated per test:
bf16 outputs carry one rounding (2^-9 relative), fp32 outputs only accumulation-order noise."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from camouflaged_vlm_b200 import check_path as cp   # noqa: E402
from camouflaged_vlm_b200 import _lib, ops          # noqa: E402

DEV = "cuda"


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _randn(*shape, seed=0, dtype=torch.float32, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


@pytest.mark.parametrize("rows,cols", [(4096, 1280), (1000, 768), (77, 256), (8, 64), (333, 1024)])
def test_layernorm(rows, cols):
    x = _randn(rows, cols, seed=1) * 3 + 0.5
    w, b = 1 + 0.1 * _randn(cols, seed=2), 0.1 * _randn(cols, seed=3)
    want = torch.nn.functional.layer_norm(x, (cols,), w, b, 1e-6)
    assert _rel(ops.layernorm(x, w, b, 1e-6, torch.float32), want) < 2e-6
    assert _rel(ops.layernorm(x, w, b, 1e-6, torch.bfloat16).float(), want) < 4e-3
    assert torch.equal(ops.layernorm(x, w, b, 1e-6, torch.bfloat16), ops.layernorm(x, w, b, 1e-6, torch.float32).bfloat16())


def test_layernorm2d_nchw():
    for batch, tokens, chans in [(2, 4096, 256), (1, 324, 256), (3, 100, 16)]:
        x = _randn(batch * tokens, chans, seed=4) * 2 - 1
        w, b = 1 + 0.1 * _randn(chans, seed=5), 0.1 * _randn(chans, seed=6)
        want = cp.layernorm2d_nchw(x, w, b, batch, tokens, 1e-6)
        assert _rel(ops.layernorm2d_nchw(x, w, b, batch, tokens, 1e-6), want) < 2e-6


def test_im2col_and_cast_are_bit_exact():
    img = _randn(2, 3, 160, 160, seed=7)
    assert torch.equal(ops.im2col_patch16(img), cp.im2col_patch16(img).bfloat16())
    x = _randn(2 * 10 * 10, 16, seed=8, dtype=torch.bfloat16)
    assert torch.equal(ops.im2col_3x3(x, 2, 10), cp.im2col_3x3(x, 2, 10))
    x = _randn(2 * 64 * 64, 256, seed=8, dtype=torch.bfloat16)
    assert torch.equal(ops.im2col_3x3(x, 2, 64), cp.im2col_3x3(x, 2, 64))
    y = _randn(1000, 64, seed=9)
    assert torch.equal(ops.cast_bf16(y), y.bfloat16())


@pytest.mark.parametrize("batch,tokens,dim,rank,rp", [(2, 100, 64, 2, 8), (1, 324, 320, 10, 16), (2, 4096, 1280, 40, 40)])
def test_adapter_embed(batch, tokens, dim, rank, rp):
    x0 = _randn(batch * tokens, dim, seed=10)
    hand = _randn(batch * tokens, rp, seed=11)
    w_t = _randn(dim, rp, seed=12, scale=0.02)
    w_t[:, rank:] = 0
    cst = _randn(tokens, rp, seed=13, scale=0.1)
    want = cp.adapter_embed(x0, hand, w_t, cst, batch, tokens, dim, rank, rp)
    got = ops.adapter_embed(x0, hand, w_t, cst, batch, tokens, dim, rank, rp).float()
    assert _rel(got, want) < 4e-3                   # one bf16 rounding of the output
    assert torch.all(got[:, rank:] == 0)


GEMM_SHAPES = [
    # m, n, k   (tile is 128 x 256 x 64)
    (128, 256, 64), (128, 256, 128), (256, 512, 256), (384, 1280, 1280), (1000, 264, 200),
    (4096, 3840, 1280), (4096, 1280, 5120), (300, 40, 768), (4096, 1280, 40),
]


@pytest.fixture(params=["1cta", "pair"])
def gemm_mode(request):
    """Run a GEMM test through both kernels: one CTA per 128x256 tile, and CTA pairs (cta_group::2) per 256x256."""
    _lib.set_option("gemm_mode", {"1cta": 1, "pair": 2}[request.param])
    yield request.param
    _lib.set_option("gemm_mode", 0)


@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
def test_gemm_plain(m, n, k, gemm_mode):
    a = _randn(m, k, seed=20, dtype=torch.bfloat16)
    b = _randn(n, k, seed=21, dtype=torch.bfloat16, scale=0.05)
    want = a.float() @ b.float().t()
    got = ops.gemm_bf16(a, b, n=n)
    assert got.shape == (m, n)
    assert _rel(got, want) < 1e-5                    # exact bf16 products, fp32 accumulation
    got16 = ops.gemm_bf16(a, b, n=n, out_dtype=torch.bfloat16)
    assert _rel(got16.float(), want) < 4e-3


def test_gemm_epilogues(gemm_mode):
    m, n, k = 512, 512, 320
    a = _randn(m, k, seed=22, dtype=torch.bfloat16)
    b = _randn(n, k, seed=23, dtype=torch.bfloat16, scale=0.05)
    bias = _randn(n, seed=24)
    res = _randn(m, n, seed=25)
    base = a.float() @ b.float().t()
    assert _rel(ops.gemm_bf16(a, b, bias=bias), base + bias) < 1e-5
    assert _rel(ops.gemm_bf16(a, b, bias=bias, gelu=True), torch.nn.functional.gelu(base + bias)) < 1e-5
    assert _rel(ops.gemm_bf16(a, b, bias=bias, residual=res), base + bias + res) < 1e-5
    # in-place residual (out aliases residual), as the proj / lin2 epilogues run it
    buf = res.clone()
    ops.gemm_bf16(a, b, bias=bias, residual=buf, out=buf)
    assert _rel(buf, base + bias + res) < 1e-5
    # row-periodic residual (pos_embed broadcast over the batch)
    pos = _randn(128, n, seed=26)
    want = base + bias + pos.repeat(m // 128, 1)
    assert _rel(ops.gemm_bf16(a, b, bias=bias, residual=pos, res_row_mod=128), want) < 1e-5


def test_gemm_two_k_segments_and_strided_operands(gemm_mode):
    m, n, k1, rp, depth = 640, 1280, 512, 40, 4
    a1 = _randn(m, k1, seed=27, dtype=torch.bfloat16)
    b1 = _randn(n, k1, seed=28, dtype=torch.bfloat16, scale=0.05)
    gall = _randn(m, depth * rp, seed=29, dtype=torch.bfloat16)
    b2 = _randn(n, rp, seed=30, dtype=torch.bfloat16, scale=0.05)
    res = _randn(m, n, seed=31)
    for i in range(depth):
        a2 = gall[:, i * rp:(i + 1) * rp]            # column slice: row stride depth*rp, 16-byte aligned base
        want = a1.float() @ b1.float().t() + a2.float() @ b2.float().t() + res
        assert _rel(ops.gemm_bf16(a1, b1, a2, b2, residual=res), want) < 1e-5
    # weights with fewer rows than n (zero rows beyond): adapter PatchEmbed2 output padded 10 -> 16
    bw = _randn(10, k1, seed=32, dtype=torch.bfloat16, scale=0.05)
    got = ops.gemm_bf16(a1, bw, n=16)
    assert _rel(got[:, :10], a1.float() @ bw.float().t()) < 1e-5
    assert torch.all(got[:, 10:] == 0)


@pytest.mark.parametrize("m,d,n", [(512, 1280, 3840), (1000, 320, 960), (4096, 768, 3072)])
def test_gemm_layernorm_fold(m, d, n, gemm_mode):
    """nn.LayerNorm folded around two GEMMs: the producer (fp32 residual epilogue) leaves bf16(x) and per-row
    partial (sum, sum of squares); the consumer normalises in its epilogue.  Reference: LN in fp32, then the GEMM."""
    k0 = 256
    a0 = _randn(m, k0, seed=60, dtype=torch.bfloat16)
    w0 = _randn(d, k0, seed=61, dtype=torch.bfloat16, scale=0.08)
    res = _randn(m, d, seed=62) * 2 + 0.7                     # a mean the fold has to subtract
    x_want = a0.float() @ w0.float().t() + res
    groups = -(-d // 128)                                     # CVLM_STAT_GROUP_COLS
    xb = torch.empty(m, d, device=DEV, dtype=torch.bfloat16)
    stat = torch.full((m, groups, 2), float("nan"), device=DEV)
    x = ops.gemm_bf16(a0, w0, residual=res, copy_bf16=xb, stat_out=stat)
    assert _rel(x, x_want) < 1e-5
    assert torch.equal(xb, x.bfloat16())
    assert _rel(stat[:, :, 0].sum(1), x.sum(1)) < 1e-5 and _rel(stat[:, :, 1].sum(1), (x * x).sum(1)) < 1e-5
    gamma, beta = 1 + 0.1 * _randn(d, seed=63), 0.1 * _randn(d, seed=64)
    w = _randn(n, d, seed=65, scale=0.05)
    b = _randn(n, seed=66, scale=0.1)
    ws = (w * gamma[None, :]).bfloat16()
    colsum = ws.double().sum(1).float()
    bias = (b.double() + w.double() @ beta.double()).float()
    want = torch.nn.functional.layer_norm(x, (d,), gamma, beta, 1e-6) @ w.t() + b
    rows = ops.ln_finalize(stat, d, 1e-6)
    mu, var = x.double().mean(1), x.double().var(1, unbiased=False)
    assert _rel(rows[:, 0], (var + 1e-6).rsqrt()) < 1e-5 and _rel(rows[:, 1], -mu * (var + 1e-6).rsqrt()) < 1e-4
    got = ops.gemm_bf16(xb, ws, bias=bias, out_dtype=torch.bfloat16, ln_rows=rows, ln_colsum=colsum)
    # bf16 x, bf16 gamma-scaled weights, bf16 output: three roundings of 2^-9
    assert _rel(got.float(), want) < 6e-3
    got_g = ops.gemm_bf16(xb, ws, bias=bias, out_dtype=torch.bfloat16, gelu=True, ln_rows=rows, ln_colsum=colsum)
    assert _rel(got_g.float(), torch.nn.functional.gelu(want)) < 8e-3


def test_gemm_rejects_bad_arguments():
    from camouflaged_vlm_b200._lib import CvlmError
    a = _randn(128, 64, dtype=torch.bfloat16)
    b = _randn(12, 64, dtype=torch.bfloat16)
    with pytest.raises(CvlmError):
        ops.gemm_bf16(a, b, n=12)                    # n % 8 != 0
    with pytest.raises(RuntimeError):
        ops.gemm_bf16(a.cpu(), b.cpu())              # no CPU fallback


ATT_CASES = [
    (2, 10, 4, 2, 32),      # head_dim 32 (tiny test config)
    # batch, grid, window, heads, head_dim
    (2, 10, 4, 2, 64),      # padded windows, generic path
    (1, 18, 14, 4, 80),     # ViT-H window size, 2x2 windows with padding
    (1, 18, 0, 4, 80),      # global, generic path (grid != 64)
    (1, 10, 14, 2, 80),     # window larger than the grid: one padded window
    (2, 64, 14, 2, 80),     # ViT-H windowed geometry (5x5 windows over 64x64)
    (1, 64, 0, 2, 80),      # ViT-H global geometry: row-aligned fast path
    (1, 64, 0, 2, 64),      # ViT-B/L global
    (1, 16, 14, 4, 64),
]


@pytest.mark.parametrize("batch,grid,window,heads,hd", ATT_CASES)
def test_attention(batch, grid, window, heads, hd):
    d = heads * hd
    s = window if window > 0 else grid
    qkv = _randn(batch * grid * grid, 3 * d, seed=40, dtype=torch.bfloat16, scale=0.8)
    relh = _randn(2 * s - 1, hd, seed=41, dtype=torch.bfloat16, scale=0.1)
    relw = _randn(2 * s - 1, hd, seed=42, dtype=torch.bfloat16, scale=0.1)
    want = cp.attention(qkv, relh, relw, batch, grid, window, heads, hd)
    got = ops.attention(qkv, relh, relw, batch, grid, window, heads, hd).float()
    # bf16 P and bf16 output: ~2^-8 relative
    assert _rel(got, want) < 8e-3


@pytest.mark.parametrize("window", [0, 14])
@pytest.mark.parametrize("qscale", [3.0, 6.0])
def test_attention_wide_logit_range(window, qscale):
    """Peaked rows: logits spread over tens of log2 units, so the single-pass softmax of the tcgen05 kernels
    has to raise its reference mid-row (the slow path) and must still match the fp32 restatement."""
    batch, grid, heads, hd = 1, 64, 2, 80
    d = heads * hd
    s = window if window > 0 else grid
    qkv = _randn(batch * grid * grid, 3 * d, seed=50, dtype=torch.bfloat16, scale=qscale)
    qkv[:, 2 * d:] = _randn(batch * grid * grid, d, seed=51, dtype=torch.bfloat16, scale=0.8)   # keep V tame
    relh = _randn(2 * s - 1, hd, seed=52, dtype=torch.bfloat16, scale=0.3)
    relw = _randn(2 * s - 1, hd, seed=53, dtype=torch.bfloat16, scale=0.3)
    want = cp.attention(qkv, relh, relw, batch, grid, window, heads, hd)
    got = ops.attention(qkv, relh, relw, batch, grid, window, heads, hd).float()
    assert torch.isfinite(got).all()
    assert _rel(got, want) < 8e-3


@pytest.mark.parametrize("batch,n", [(2, 160), (1, 256), (2, 288), (2, 1024)])
def test_high_pass_native(batch, n):
    """PromptGenerator.fft as two tcgen05 GEMMs with the bf16 low-pass projector vs the fp32 FFT route.
    Two bf16 contractions of length n: ~2e-3 relative (the result is rounded to bf16 by the next kernel anyway)."""
    from camouflaged_vlm_b200.packing import high_pass_matrices
    x = _randn(batch, 3, n, n, seed=70)
    pcat, npr, pi = high_pass_matrices(n, 0.25, DEV)
    got = ops.high_pass(x, pcat, npr, pi)
    want = cp.high_pass(x, 0.25)
    assert got.shape == want.shape and torch.all(got >= 0)
    assert _rel(got, want) < 5e-3
    # images are independent: a batch of one gives the same bits
    assert torch.equal(ops.high_pass(x[:1].contiguous(), pcat, npr, pi), got[:1])


def test_gemm_random_shapes(gemm_mode):
    """Seeded sweep over ragged shapes (every M / N / K edge case of the tile, pair and batch logic) with random
    epilogue combinations, against fp32 torch."""
    import random
    rnd = random.Random(1234)
    for it in range(24):
        m = rnd.choice([1, 7, 127, 128, 129, 255, 256, 257, 300, 513, 1000, 2049])
        n = 8 * rnd.randint(1, 90)
        k = 8 * rnd.randint(1, 100)
        a = _randn(m, k, seed=100 + it, dtype=torch.bfloat16)
        b = _randn(n, k, seed=200 + it, dtype=torch.bfloat16, scale=0.05)
        bias = _randn(n, seed=300 + it) if rnd.random() < 0.7 else None
        res = _randn(m, n, seed=400 + it) if rnd.random() < 0.5 else None
        gelu = rnd.random() < 0.3
        want = a.float() @ b.float().t()
        if bias is not None:
            want = want + bias
        if gelu:
            want = torch.nn.functional.gelu(want)
        if res is not None:
            want = want + res
        got = ops.gemm_bf16(a, b, bias=bias, residual=res, gelu=gelu)
        assert _rel(got, want) < 2e-5, (it, m, n, k, gelu)
        got16 = ops.gemm_bf16(a, b, bias=bias, residual=res, gelu=gelu, out_dtype=torch.bfloat16)
        assert _rel(got16.float(), want) < 4e-3, (it, m, n, k, gelu)


@pytest.mark.parametrize("batch,tokens,heads,hd", [(2, 579, 16, 64), (1, 64, 2, 64), (3, 65, 2, 80), (2, 256, 1, 64),
                                                   (1, 257, 3, 80), (2, 1, 2, 64), (1, 1000, 2, 64)])
def test_attention_plain(batch, tokens, heads, hd):
    """The global-attention pipeline without relative-position terms over an arbitrary sequence length (ragged last key
    unit masked, query rows beyond the sequence not stored) against fp32 softmax attention on the same bf16 inputs.
    579 tokens x 16 heads x 64 is the Alpha-CLIP ViT-L/14@336 tower with two MaPLe context tokens."""
    d = heads * hd
    qkv = _randn(batch * tokens, 3 * d, seed=90 + tokens, dtype=torch.bfloat16, scale=0.8)
    q, k, v = (t.reshape(batch, tokens, heads, hd).permute(0, 2, 1, 3) for t in qkv.float().split(d, dim=1))
    want = (torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, dim=-1) @ v).permute(0, 2, 1, 3).reshape(batch * tokens, d)
    got = ops.attention_plain(qkv, batch, tokens, heads, hd).float()
    assert torch.isfinite(got).all()
    assert _rel(got, want) < 8e-3
    # the last query block is ragged: nothing may be written past the last token
    from camouflaged_vlm_b200 import _lib
    out = torch.full((batch * tokens + 256, d), 7.0, device=DEV, dtype=torch.bfloat16)
    _lib.check(_lib.load().cvlm_op_attention_plain(qkv.data_ptr(), out.data_ptr(), batch, tokens, heads, hd, _lib.stream_ptr()))
    assert torch.equal(out[:batch * tokens].float(), got) and bool((out[batch * tokens:] == 7.0).all())


@pytest.mark.parametrize("m,n,k", [(579, 1024, 1024), (579, 1024, 4096), (579, 3072, 1024), (1737, 1024, 1024), (130, 520, 200),
                                   (64, 4096, 1024), (1, 776, 64)])
def test_gemm_few_tiles_narrow_n(m, n, k):
    """Small-M problems (the batch-1 latency path) are cut into 128- or 64-column tiles by the launcher so that more SMs
    work; every epilogue must give the same numbers as the 256-column kernels, bit for bit (same K order per element)."""
    _lib.set_option("gemm_mode", 0)
    a = _randn(m, k, seed=500 + m, dtype=torch.bfloat16)
    b = _randn(n, k, seed=600 + n, dtype=torch.bfloat16, scale=0.05)
    bias = _randn(n, seed=700 + k)
    res = _randn(m, n, seed=800 + k)
    want = a.float() @ b.float().t() + bias
    got = ops.gemm_bf16(a, b, bias=bias, residual=res)
    assert _rel(got, want + res) < 2e-5
    got16 = ops.gemm_bf16(a, b, bias=bias, gelu=True, out_dtype=torch.bfloat16)
    assert _rel(got16.float(), torch.nn.functional.gelu(want)) < 4e-3
    _lib.set_option("gemm_mode", 1)                       # the 256-column kernel on the same problem
    try:
        assert torch.equal(ops.gemm_bf16(a, b, bias=bias, residual=res), got)
        assert torch.equal(ops.gemm_bf16(a, b, bias=bias, gelu=True, out_dtype=torch.bfloat16), got16)
    finally:
        _lib.set_option("gemm_mode", 0)
