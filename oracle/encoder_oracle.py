"""CPU restatement of the reference SAM-Adapter ``ImageEncoderViT`` forward.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this file; nothing under
``camouflaged_vlm_b200/`` does, and the product path raises when its CUDA library is missing.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in
the build container by ``oracle/make_golden.py`` (which imports the unmodified module from
/root/reference) and committed under ``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks
the oracle against them on every run, and against the live reference when that tree is present.

The algorithm is written here in functional form over a plain ``state_dict`` (no nn.Module),
each step citing the reference lines it restates.  All ``file:line`` references are to
``/root/reference/models/mmseg/models/sam/``.
"""
import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def high_pass(x: Tensor, rate: float) -> Tensor:
    """PromptGenerator.fft, image_encoder.py:332-353.

    fft2(norm=forward) -> fftshift over ALL dims -> zero the centred (2*line)^2 box -> ifftshift ->
    ifft2(norm=forward).real.abs().  The all-dims shift only permutes batch/channel and is undone
    by the matching ifftshift; the mask does not depend on batch/channel, so shifting the two
    spatial dims is equivalent (checked against the reference in tests).
    """
    w, h = x.shape[-2:]
    line = int((w * h * rate) ** 0.5 // 2)
    mask = torch.zeros_like(x)
    mask[..., w // 2 - line:w // 2 + line, h // 2 - line:h // 2 + line] = 1
    spec = torch.fft.fftshift(torch.fft.fft2(x, norm="forward"), dim=(-2, -1))
    spec = spec * (1 - mask)
    inv = torch.fft.ifft2(torch.fft.ifftshift(spec, dim=(-2, -1)), norm="forward").real
    return inv.abs()


def rel_pos_table(q_size: int, k_size: int, rel_pos: Tensor) -> Tensor:
    """get_rel_pos, image_encoder.py:556-586: R[i, j, :] = rel_pos[i - j + (k_size - 1)] when
    q_size == k_size and the table already has 2*S-1 rows; otherwise linear interpolation."""
    max_rel = int(2 * max(q_size, k_size) - 1)
    if rel_pos.shape[0] != max_rel:
        rp = F.interpolate(rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1),
                           size=max_rel, mode="linear")
        rp = rp.reshape(-1, max_rel).permute(1, 0)
    else:
        rp = rel_pos
    qc = torch.arange(q_size)[:, None] * max(k_size / q_size, 1.0)
    kc = torch.arange(k_size)[None, :] * max(q_size / k_size, 1.0)
    idx = (qc - kc) + (k_size - 1) * max(q_size / k_size, 1.0)
    return rp[idx.long()]


def attention(x: Tensor, sd: Dict[str, Tensor], pre: str, heads: int, use_rel_pos: bool) -> Tensor:
    """Attention.forward, image_encoder.py:488-504, on x[B', H, W, D] (a window batch or whole grids).
    No mask: zero-padded window tokens take part as keys.  Rel-pos uses the UNSCALED q
    (image_encoder.py:495 vs 498; add_decomposed_rel_pos :589-625)."""
    b, h, w, d = x.shape
    hd = d // heads
    qkv = F.linear(x, sd[pre + "qkv.weight"], sd.get(pre + "qkv.bias"))
    qkv = qkv.reshape(b, h * w, 3, heads, hd).permute(2, 0, 3, 1, 4).reshape(3, b * heads, h * w, hd)
    q, k, v = qkv[0], qkv[1], qkv[2]
    logits = (q * hd ** -0.5) @ k.transpose(-2, -1)
    if use_rel_pos:
        rh = rel_pos_table(h, h, sd[pre + "rel_pos_h"])
        rw = rel_pos_table(w, w, sd[pre + "rel_pos_w"])
        rq = q.reshape(b * heads, h, w, hd)
        rel_h = torch.einsum("bhwc,hkc->bhwk", rq, rh)
        rel_w = torch.einsum("bhwc,wkc->bhwk", rq, rw)
        logits = (logits.view(b * heads, h, w, h, w) + rel_h[..., :, None] + rel_w[..., None, :]
                  ).view(b * heads, h * w, h * w)
    p = logits.softmax(dim=-1)
    o = (p @ v).view(b, heads, h, w, hd).permute(0, 2, 3, 1, 4).reshape(b, h, w, d)
    return F.linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def window_partition(x: Tensor, ws: int) -> Tuple[Tensor, Tuple[int, int]]:
    """image_encoder.py:507-528: zero-pad bottom/right to a multiple of ws, split into windows."""
    b, h, w, c = x.shape
    ph, pw = (ws - h % ws) % ws, (ws - w % ws) % ws
    if ph or pw:
        x = F.pad(x, (0, 0, 0, pw, 0, ph))
    hp, wp = h + ph, w + pw
    x = x.view(b, hp // ws, ws, wp // ws, ws, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, ws, ws, c), (hp, wp)


def window_unpartition(win: Tensor, ws: int, pad_hw: Tuple[int, int], hw: Tuple[int, int]) -> Tensor:
    """image_encoder.py:531-553: inverse of window_partition, cropping the padding."""
    hp, wp = pad_hw
    h, w = hw
    b = win.shape[0] // (hp * wp // ws // ws)
    x = win.view(b, hp // ws, wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(b, hp, wp, -1)
    return x[:, :h, :w, :]


def block(x: Tensor, sd: Dict[str, Tensor], pre: str, heads: int, ws: int, use_rel_pos: bool,
          eps: float) -> Tensor:
    """Block.forward, image_encoder.py:430-446 + MLPBlock common.py:25-26 (GELU = exact erf)."""
    d = x.shape[-1]
    y = F.layer_norm(x, (d,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps)
    if ws > 0:
        hw = (y.shape[1], y.shape[2])
        y, pad_hw = window_partition(y, ws)
    y = attention(y, sd, pre + "attn.", heads, use_rel_pos)
    if ws > 0:
        y = window_unpartition(y, ws, pad_hw, hw)
    x = x + y
    y = F.layer_norm(x, (d,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps)
    y = F.linear(F.gelu(F.linear(y, sd[pre + "mlp.lin1.weight"], sd[pre + "mlp.lin1.bias"])),
                 sd[pre + "mlp.lin2.weight"], sd[pre + "mlp.lin2.bias"])
    return x + y


def layer_norm_2d(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """LayerNorm2d, common.py:31-43: biased variance over the channel dim of NCHW."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return w[:, None, None] * ((x - u) / torch.sqrt(s + eps)) + b[:, None, None]


def prompts(sd: Dict[str, Tensor], cfg, x_patch: Tensor, inp: Tensor) -> List[Tensor]:
    """PromptGenerator.init_embeddings / init_handcrafted / get_prompt, image_encoder.py:278-296.
    x_patch is the patch embedding BEFORE pos_embed (image_encoder.py:134-138)."""
    pg = "prompt_generator."
    b = x_patch.shape[0]
    # init_embeddings (:278-281) takes N,C,H,W from a permuted VIEW but then reshapes the NHWC
    # tensor itself to (N, C, H*W): the flat NHWC buffer of each image is re-read as a [C, H*W]
    # matrix and transposed, so "token t, feature c" is flat[c * H*W + t] -- a scramble of tokens
    # and channels, not the patch embedding of token t.  Reproduced as is.
    scr = x_patch.reshape(b, cfg.embed_dim, -1).permute(0, 2, 1)
    emb = F.linear(scr, sd[pg + "embedding_generator.weight"], sd[pg + "embedding_generator.bias"])
    hand = F.conv2d(high_pass(inp, cfg.freq_nums), sd[pg + "prompt_generator.proj.weight"],
                    sd[pg + "prompt_generator.proj.bias"], stride=cfg.patch_size)
    hand = hand.reshape(b, cfg.rank, -1).permute(0, 2, 1)
    out = []
    for i in range(cfg.depth):
        t = F.gelu(F.linear(hand + emb, sd[pg + f"lightweight_mlp_{i}.0.weight"],
                            sd[pg + f"lightweight_mlp_{i}.0.bias"]))
        out.append(F.linear(t, sd[pg + "shared_mlp.weight"], sd[pg + "shared_mlp.bias"]))
    return out


def encoder_forward_autograd(sd: Dict[str, Tensor], cfg, x: Tensor, interm: bool = False, ln_eps: float = 1e-6,
                             stop_after_block: int = -1):
    """ImageEncoderViT.forward, image_encoder.py:132-155.  ``ln_eps`` is the eps the callers bind
    into norm_layer (models/sam_maskdecoder_edge.py:129).  ``stop_after_block`` >= 0 returns the
    NHWC residual stream after that block (debug aid for per-layer parity).  Plain torch ops: differentiable
    (``adapter_gradients``); ``encoder_forward`` is the same function under ``torch.no_grad()``."""
    inp = x
    g = cfg.grid
    xp = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"],
                  stride=cfg.patch_size).permute(0, 2, 3, 1)          # PatchEmbed :655-659
    pr = prompts(sd, cfg, xp, inp)
    x = xp + sd["pos_embed"] if cfg.use_abs_pos else xp
    b = x.shape[0]
    interms = []
    for i in range(cfg.depth):
        x = pr[i].reshape(b, g, g, -1) + x                            # :145
        x = block(x, sd, f"blocks.{i}.", cfg.num_heads, cfg.block_window(i), cfg.use_rel_pos, ln_eps)
        if i == stop_after_block:
            return x
        if interm and cfg.block_window(i) == 0:
            interms.append(x.permute(0, 3, 1, 2))                     # :147-148
    y = F.conv2d(x.permute(0, 3, 1, 2), sd["neck.0.weight"])          # neck :98-114
    y = layer_norm_2d(y, sd["neck.1.weight"], sd["neck.1.bias"])
    y = F.conv2d(y, sd["neck.2.weight"], padding=1)
    y = layer_norm_2d(y, sd["neck.3.weight"], sd["neck.3.bias"])
    return (y, interms) if interm else y


encoder_forward = torch.no_grad()(encoder_forward_autograd)


def cotangents(cfg, batch: int, seed: int = 0) -> Tuple[Tensor, List[Tensor]]:
    """Seeded d(loss)/d(output) tensors for gradient parity: loss = sum(out * c) + sum_k sum(interm_k * c_k)."""
    from camouflaged_vlm_b200.synth import _normal
    g = cfg.grid
    c_out = torch.from_numpy(_normal("cotangent.out", (batch, cfg.out_chans, g, g), seed))
    c_int = [torch.from_numpy(_normal(f"cotangent.interm{k}", (batch, cfg.embed_dim, g, g), seed)) * 0.05
             for k in range(len(cfg.global_attn_indexes))]
    return c_out, c_int


def adapter_gradients(sd: Dict[str, Tensor], cfg, x: Tensor, c_out: Tensor, c_interm: List[Tensor] = None,
                      ln_eps: float = 1e-6) -> Dict[str, Tensor]:
    """Gradients of  sum(out * c_out) [+ sum_k sum(interm_k * c_interm[k])]  with respect to every
    ``prompt_generator.*`` parameter, everything else frozen -- the training configuration of the reference
    (train_ovcos_maskdecoder_edge.py:301-303: only names without "image_encoder" or with "prompt_generator" train;
    the gradient reaches the adapter through all blocks, models/sam_maskdecoder_edge.py:429-460)."""
    leaf = {k: (v.detach().clone().requires_grad_(True) if k.startswith("prompt_generator.") else v.detach())
            for k, v in sd.items()}
    use_interm = bool(c_interm)
    res = encoder_forward_autograd(leaf, cfg, x, interm=use_interm, ln_eps=ln_eps)
    out, interms = res if use_interm else (res, [])
    loss = (out * c_out).sum()
    for t, c in zip(interms, c_interm or []):
        loss = loss + (t * c).sum()
    loss.backward()
    return {k: v.grad for k, v in leaf.items() if k.startswith("prompt_generator.")}


def rel_l2(a: Tensor, ref: Tensor) -> float:
    a, ref = a.double().flatten(), ref.double().flatten()
    return float((a - ref).norm() / ref.norm().clamp_min(1e-30))
