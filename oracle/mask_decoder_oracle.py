"""CPU restatement of the downstream probe: MaskDecoder_Edge + TwoWayTransformer_MaskDecoder_Edge +
PositionEmbeddingRandom + postprocess (reference: models/mmseg/models/sam/mask_decoder_edge.py:96-190,
transformer_maskdecoder_edge.py:62-272, models/sam_maskdecoder_edge.py:85-110, 347-388).

TEST INFRASTRUCTURE -- NOT PRODUCT CODE (same rules as encoder_oracle.py).  It exists only to turn an
image embedding into a binary mask so that the encoder's parity can be stated downstream
(north_star: mask IoU >= 0.99).  Pinned against the reference decoder's own output in
tests/golden/decoder_*.npz (oracle/make_golden.py).  Functional form over a state_dict whose keys are
the reference module's keys; the decoder only works at batch 1, so callers loop.
"""
import math
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from camouflaged_vlm_b200.synth import _normal

Tensor = torch.Tensor
DIM, HEADS, MLP_DIM, DEPTH = 256, 8, 2048, 2      # models/sam_maskdecoder_edge.py:137-148
NUM_MASK_TOKENS = 4                                # num_multimask_outputs 3 + 1 (mask_decoder_edge.py:50)


def decoder_param_shapes() -> Dict[str, Tuple[int, ...]]:
    d, q = DIM, DIM // 4
    e = DIM // 8
    s: Dict[str, Tuple[int, ...]] = {}

    def lin(name, i, o):
        s[name + ".weight"] = (o, i)
        s[name + ".bias"] = (o,)

    def attn(name, down):
        inner = d // down
        for p in ("q_proj", "k_proj", "v_proj"):
            lin(f"{name}.{p}", d, inner)
        lin(f"{name}.out_proj", inner, d)

    def norm(name, c=d):
        s[name + ".weight"] = (c,)
        s[name + ".bias"] = (c,)

    for i in range(DEPTH):
        b = f"transformer.layers.{i}."
        attn(b + "self_attn", 1)
        norm(b + "norm1")
        attn(b + "cross_attn_token_to_image", 2)
        norm(b + "norm2")
        attn(b + "cross_attn_token_to_cond", 2)
        norm(b + "norm2_cond")
        lin(b + "mlp.lin1", d, MLP_DIM)
        lin(b + "mlp.lin2", MLP_DIM, d)
        norm(b + "norm3")
        norm(b + "norm4")
        attn(b + "cross_attn_image_to_token", 2)
        norm(b + "norm4_cond")
        attn(b + "cross_attn_image_to_cond", 2)
    attn("transformer.final_attn_token_to_image", 2)
    norm("transformer.norm_final_attn")
    s["iou_token.weight"] = (1, d)
    s["mask_tokens.weight"] = (NUM_MASK_TOKENS, d)
    s["output_upscaling.0.weight"] = (d, q, 2, 2)
    s["output_upscaling.0.bias"] = (q,)
    norm("output_upscaling.1", q)
    s["output_upscaling.3.weight"] = (q, e, 2, 2)
    s["output_upscaling.3.bias"] = (e,)
    for i in range(NUM_MASK_TOKENS):
        lin(f"output_hypernetworks_mlps.{i}.layers.0", d, d)
        lin(f"output_hypernetworks_mlps.{i}.layers.1", d, d)
        lin(f"output_hypernetworks_mlps.{i}.layers.2", d, e)
    lin("iou_prediction_head.layers.0", d, 256)
    lin("iou_prediction_head.layers.1", 256, 256)
    lin("iou_prediction_head.layers.2", 256, NUM_MASK_TOKENS)
    s["edge_token.weight"] = (1, d)
    lin("edge_mlp.layers.0", d, d)
    lin("edge_mlp.layers.1", d, d)
    lin("edge_mlp.layers.2", d, e)
    s["embedding_encoder.0.weight"] = (d, q, 2, 2)
    s["embedding_encoder.0.bias"] = (q,)
    norm("embedding_encoder.1", q)
    s["embedding_encoder.3.weight"] = (q, e, 2, 2)
    s["embedding_encoder.3.bias"] = (e,)
    s["embedding_maskfeature.0.weight"] = (e, q, 3, 3)
    s["embedding_maskfeature.0.bias"] = (q,)
    norm("embedding_maskfeature.1", q)
    s["embedding_maskfeature.3.weight"] = (q, e, 3, 3)
    s["embedding_maskfeature.3.bias"] = (e,)
    return s


def seeded_decoder_state(seed: int = 0, std: float = 0.05) -> Dict[str, Tensor]:
    """Deterministic decoder weights (+ the probe's prompt inputs) keyed by name, like synth.py."""
    sd = {}
    for name, shape in decoder_param_shapes().items():
        v = _normal("decoder." + name, shape, seed) * np.float32(std)
        is_norm_w = name.endswith(".weight") and len(shape) == 1
        sd[name] = torch.from_numpy(v + np.float32(1.0) if is_norm_w else v)
    sd["probe.pe_gaussian"] = torch.from_numpy(_normal("decoder.pe", (2, DIM // 2), seed))      # PositionEmbeddingRandom
    sd["probe.sparse"] = torch.from_numpy(_normal("decoder.sparse", (1, 2, DIM), seed))         # [img_feat, txt_feat] stand-in
    sd["probe.no_mask_embed"] = torch.from_numpy(_normal("decoder.nomask", (DIM,), seed) * np.float32(std))
    return sd


def _attention(sd, pre, q, k, v):
    """transformer_maskdecoder_edge.py Attention.forward."""
    def proj(n, x):
        return F.linear(x, sd[f"{pre}.{n}.weight"], sd[f"{pre}.{n}.bias"])

    def heads(x):
        b, n, c = x.shape
        return x.reshape(b, n, HEADS, c // HEADS).transpose(1, 2)

    q, k, v = heads(proj("q_proj", q)), heads(proj("k_proj", k)), heads(proj("v_proj", v))
    a = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(q.shape[-1]), dim=-1) @ v
    b, h, n, c = a.shape
    return proj("out_proj", a.transpose(1, 2).reshape(b, n, h * c))


def _ln(sd, name, x):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def _ln2d(sd, name, x, eps=1e-6):
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return sd[name + ".weight"][:, None, None] * ((x - u) / torch.sqrt(s + eps)) + sd[name + ".bias"][:, None, None]


def _mlp3(sd, pre, x):
    x = F.relu(F.linear(x, sd[pre + ".layers.0.weight"], sd[pre + ".layers.0.bias"]))
    x = F.relu(F.linear(x, sd[pre + ".layers.1.weight"], sd[pre + ".layers.1.bias"]))
    return F.linear(x, sd[pre + ".layers.2.weight"], sd[pre + ".layers.2.bias"])


def _two_way_transformer(sd, src, pos, tokens, cond):
    """TwoWayTransformer.forward + TwoWayAttentionBlock.forward (transformer_maskdecoder_edge.py:62-214)."""
    keys = src.flatten(2).permute(0, 2, 1)
    key_pe = pos.flatten(2).permute(0, 2, 1)
    queries, query_pe = tokens, tokens
    for i in range(DEPTH):
        b = f"transformer.layers.{i}."
        if i == 0:      # skip_first_layer_pe
            queries = _attention(sd, b + "self_attn", queries, queries, queries)
        else:
            q = queries + query_pe
            queries = queries + _attention(sd, b + "self_attn", q, q, queries)
        queries = _ln(sd, b + "norm1", queries)
        queries = _ln(sd, b + "norm2", queries + _attention(sd, b + "cross_attn_token_to_image",
                                                             queries + query_pe, keys + key_pe, keys))
        queries = _ln(sd, b + "norm2_cond", queries + _attention(sd, b + "cross_attn_token_to_cond",
                                                                  queries + query_pe, cond + cond, cond))
        m = F.linear(F.relu(F.linear(queries, sd[b + "mlp.lin1.weight"], sd[b + "mlp.lin1.bias"])),
                     sd[b + "mlp.lin2.weight"], sd[b + "mlp.lin2.bias"])
        queries = _ln(sd, b + "norm3", queries + m)
        keys = _ln(sd, b + "norm4_cond", keys + _attention(sd, b + "cross_attn_image_to_cond",
                                                            keys + key_pe, cond + cond, cond))
        keys = _ln(sd, b + "norm4", keys + _attention(sd, b + "cross_attn_image_to_token",
                                                       keys + key_pe, queries + query_pe, queries))
    a = _attention(sd, "transformer.final_attn_token_to_image", queries + query_pe, keys + key_pe, keys)
    return _ln(sd, "transformer.norm_final_attn", queries + a), keys


def dense_pe(sd, size: int) -> Tensor:
    """PositionEmbeddingRandom.forward(size), models/sam_maskdecoder_edge.py:99-110."""
    grid = torch.ones((size, size), dtype=torch.float32)
    y = (grid.cumsum(0) - 0.5) / size
    x = (grid.cumsum(1) - 0.5) / size
    c = (2 * torch.stack([x, y], dim=-1) - 1) @ sd["probe.pe_gaussian"]
    c = 2 * np.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1).permute(2, 0, 1)[None]


@torch.no_grad()
def decode_all(sd: Dict[str, Tensor], emb: Tensor, sparse: Tensor = None) -> Dict[str, Tensor]:
    """mask_decoder_edge.py forward/predict_masks on ONE image embedding [1,256,g,g] with the prompt `sparse` [1,n,256]
    (default: the probe's).  Returns every output of predict_masks (:140-190): masks [1,4,4g,4g] (all mask tokens, after
    the edge modulation), edge [1,1,4g,4g], iou [1,4], plus the transformer's output tokens hs [1,6,256]."""
    emb = emb.float().cpu()
    sparse = sd["probe.sparse"] if sparse is None else sparse.float().cpu()
    g = emb.shape[-1]
    up = lambda pre, x: F.conv_transpose2d(x, sd[pre + ".weight"], sd[pre + ".bias"], stride=2)
    edge_feat = up("embedding_encoder.3", F.gelu(_ln2d(sd, "embedding_encoder.1", up("embedding_encoder.0", emb))))
    tokens = torch.cat([sd["iou_token.weight"], sd["mask_tokens.weight"], sd["edge_token.weight"]], 0)[None]
    src = emb + sd["probe.no_mask_embed"].reshape(1, -1, 1, 1).expand(1, -1, g, g)
    hs, keys = _two_way_transformer(sd, src, dense_pe(sd, g), tokens, sparse)
    mask_tokens_out = hs[:, 1:1 + NUM_MASK_TOKENS + 1, :]
    src2 = keys.transpose(1, 2).view(1, DIM, g, g)
    upscaled = F.gelu(up("output_upscaling.3", F.gelu(_ln2d(sd, "output_upscaling.1", up("output_upscaling.0", src2)))))
    ct3 = lambda pre, x: F.conv_transpose2d(x, sd[pre + ".weight"], sd[pre + ".bias"], stride=1, padding=1)
    edge_emb = ct3("embedding_maskfeature.3", F.gelu(_ln2d(sd, "embedding_maskfeature.1",
                                                           ct3("embedding_maskfeature.0", upscaled)))) + edge_feat
    hyper = [_mlp3(sd, f"output_hypernetworks_mlps.{i}", mask_tokens_out[:, i]) for i in range(NUM_MASK_TOKENS)]
    hyper.append(_mlp3(sd, "edge_mlp", mask_tokens_out[:, NUM_MASK_TOKENS]))
    hyper = torch.stack(hyper, 1)
    b, c, h, w = upscaled.shape
    masks = (hyper[:, :NUM_MASK_TOKENS] @ upscaled.view(b, c, h * w)).view(b, -1, h, w)
    edge = torch.sigmoid((hyper[:, NUM_MASK_TOKENS:] @ edge_emb.view(b, c, h * w)).view(b, -1, h, w))
    masks = masks * edge + masks
    return {"masks": masks, "edge": edge, "iou": _mlp3(sd, "iou_prediction_head", hs[:, 0]), "hs": hs,
            "keys": keys, "upscaled": upscaled, "edge_emb": edge_emb, "hyper": hyper}


def decode_mask_logits(sd: Dict[str, Tensor], emb: Tensor) -> Tensor:
    """multimask_output=False: low-res mask logits [1,1,4g,4g] (the caller upsamples and thresholds at 0,
    models/sam_maskdecoder_edge.py:355, 359-388)."""
    return decode_all(sd, emb)["masks"][:, 0:1]


def binary_mask(logits: Tensor, out_size: int) -> Tensor:
    """postprocess_masks (bilinear to the input size) then > 0 (models/sam_maskdecoder_edge.py:359-388)."""
    return F.interpolate(logits, (out_size, out_size), mode="bilinear", align_corners=False) > 0


def iou(a: Tensor, b: Tensor) -> float:
    inter = (a & b).sum().item()
    union = (a | b).sum().item()
    return inter / union if union else 1.0
