"""Drop-in replacement for the reference mask decoder:

* ``MaskDecoder`` -- ``MaskDecoder`` of models/mmseg/models/sam/mask_decoder_edge.py:16-190 (imported by the
  cascade as ``MaskDecoder_Edge``): same constructor, same parameter tree / state_dict keys, same
  ``forward(image_embeddings, interm_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings,
  multimask_output) -> (masks, edge, iou_pred)``;
* ``TwoWayTransformer`` -- parameter tree of transformer_maskdecoder_edge.py:16-60, 112-157 (the decoder accepts the
  reference's own transformer instance as well: only its parameters are read);
* ``postprocess_masks`` -- models/sam_maskdecoder_edge.py:359-388.

The forward runs csrc/decoder.cu through the C ABI (include/cvlm.h ``cvlm_decoder_*``).  Unlike the reference, which
only works at batch 1 (it repeat-interleaves the embeddings by the prompt batch), image b is decoded with prompt b, so a
batch of embeddings straight out of the encoder is decoded in one call.  There is no CPU path and no PyTorch fallback.
The sub-modules are parameter containers; they carry no compute.
"""
import ctypes as C
from typing import Dict, List, Optional, Tuple, Type

import torch
import torch.nn as nn

from . import _lib
from .common import LayerNorm2d, MLPBlock

NUM_TOKENS = 6           # iou token + 4 mask tokens + edge token


class Attention(nn.Module):
    """Parameter container for transformer_maskdecoder_edge.py:217-272."""

    def __init__(self, embedding_dim: int, num_heads: int, downsample_rate: int = 1) -> None:
        super().__init__()
        self.embedding_dim = embedding_dim
        self.internal_dim = embedding_dim // downsample_rate
        self.num_heads = num_heads
        assert self.internal_dim % num_heads == 0, "num_heads must divide embedding_dim."
        self.q_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.k_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.v_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.out_proj = nn.Linear(self.internal_dim, embedding_dim)


class TwoWayAttentionBlock(nn.Module):
    """Parameter container for transformer_maskdecoder_edge.py:112-157."""

    def __init__(self, embedding_dim: int, num_heads: int, mlp_dim: int = 2048, activation: Type[nn.Module] = nn.ReLU,
                 attention_downsample_rate: int = 2, skip_first_layer_pe: bool = False) -> None:
        super().__init__()
        self.self_attn = Attention(embedding_dim, num_heads)
        self.norm1 = nn.LayerNorm(embedding_dim)
        self.cross_attn_token_to_image = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm2 = nn.LayerNorm(embedding_dim)
        self.cross_attn_token_to_cond = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm2_cond = nn.LayerNorm(embedding_dim)
        self.mlp = MLPBlock(embedding_dim, mlp_dim, activation)
        self.norm3 = nn.LayerNorm(embedding_dim)
        self.norm4 = nn.LayerNorm(embedding_dim)
        self.cross_attn_image_to_token = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm4_cond = nn.LayerNorm(embedding_dim)
        self.cross_attn_image_to_cond = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.skip_first_layer_pe = skip_first_layer_pe


class TwoWayTransformer(nn.Module):
    """Parameter container for transformer_maskdecoder_edge.py:16-60."""

    def __init__(self, depth: int, embedding_dim: int, num_heads: int, mlp_dim: int,
                 activation: Type[nn.Module] = nn.ReLU, attention_downsample_rate: int = 2) -> None:
        super().__init__()
        if activation is not nn.ReLU or attention_downsample_rate != 2:
            raise NotImplementedError("the B200 decoder implements the reference's ReLU MLP and downsample rate 2")
        self.depth = depth
        self.embedding_dim = embedding_dim
        self.num_heads = num_heads
        self.mlp_dim = mlp_dim
        self.layers = nn.ModuleList(
            TwoWayAttentionBlock(embedding_dim, num_heads, mlp_dim, activation, attention_downsample_rate,
                                 skip_first_layer_pe=(i == 0)) for i in range(depth))
        self.final_attn_token_to_image = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm_final_attn = nn.LayerNorm(embedding_dim)


class MLP(nn.Module):
    """Parameter container for mask_decoder_edge.py:195-217."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, num_layers: int, sigmoid_output: bool = False):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))
        self.sigmoid_output = sigmoid_output


def pack_decoder_weights(sd: Dict[str, torch.Tensor], depth: int, device) -> Dict[str, torch.Tensor]:
    """Device tensors registered with ``cvlm_decoder_set_tensor``:

    * every fp32 parameter under its reference name, and every nn.Linear weight also transposed as ``<name>.weight.t``
      ([in, out]: what the token-side kernels stream);
    * ``tokens`` fp32 [6,256] = [iou_token; mask_tokens; edge_token] (mask_decoder_edge.py:157);
    * ``<attention>.{k,v,q,out}_proj.weight.bf16`` for the projections applied to the 4096 image tokens;
    * k2 s2 transposed convolutions as GEMMs: ``<conv>.wp`` bf16 [(dy*2+dx)*Cout + o, i] = W[i,o,dy,dx] and ``<conv>.bp``
      fp32 = bias tiled 4x;
    * 3x3 s1 p1 transposed convolutions as implicit GEMMs (3x3 convolutions with the flipped kernel):
      ``<conv>.wp`` bf16 [o, (ty*3+tx)*64 + i] = W[i,o,2-ty,2-tx], 64 input channels per tap (zero padded).
    """
    out: Dict[str, torch.Tensor] = {}
    f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
    b16 = lambda t: t.detach().to(device=device, dtype=torch.float32).to(torch.bfloat16).contiguous()
    for k, v in sd.items():
        out[k] = f32(v)
        if v.dim() == 2 and k.endswith(".weight") and k not in ("iou_token.weight", "mask_tokens.weight", "edge_token.weight"):
            out[k + ".t"] = out[k].t().contiguous()          # nn.Linear weights, transposed for the token-side kernels
    out["tokens"] = torch.cat([out["iou_token.weight"], out["mask_tokens.weight"], out["edge_token.weight"]], 0).contiguous()
    img_side = []
    for i in range(depth):
        p = f"transformer.layers.{i}."
        img_side += [p + "cross_attn_token_to_image.k_proj", p + "cross_attn_token_to_image.v_proj",
                     p + "cross_attn_image_to_cond.q_proj", p + "cross_attn_image_to_cond.out_proj",
                     p + "cross_attn_image_to_token.q_proj", p + "cross_attn_image_to_token.out_proj"]
    img_side += ["transformer.final_attn_token_to_image.k_proj", "transformer.final_attn_token_to_image.v_proj"]
    for n in img_side:
        out[n + ".weight.bf16"] = b16(sd[n + ".weight"])
    for n in ("output_upscaling.0", "output_upscaling.3", "embedding_encoder.0", "embedding_encoder.3"):
        w = f32(sd[n + ".weight"])                      # [Cin, Cout, 2, 2]
        cin, cout = w.shape[0], w.shape[1]
        out[n + ".wp"] = w.permute(2, 3, 1, 0).reshape(4 * cout, cin).to(torch.bfloat16).contiguous()
        out[n + ".bp"] = f32(sd[n + ".bias"]).repeat(4).contiguous()
    for n in ("embedding_maskfeature.0", "embedding_maskfeature.3"):
        w = f32(sd[n + ".weight"])                      # [Cin, Cout, 3, 3]
        cin, cout = w.shape[0], w.shape[1]
        full = torch.zeros(cout, 9, 64, device=device, dtype=torch.float32)      # 64 input channels per tap (zero padded)
        full[:, :, :cin] = w.flip(2, 3).permute(1, 2, 3, 0).reshape(cout, 9, cin)
        out[n + ".wp"] = full.reshape(cout, 9 * 64).to(torch.bfloat16).contiguous()
    return out


class MaskDecoder(nn.Module):
    def __init__(self, *, transformer_dim: int, transformer: nn.Module, num_multimask_outputs: int = 3,
                 activation: Type[nn.Module] = nn.GELU, iou_head_depth: int = 3, iou_head_hidden_dim: int = 256) -> None:
        super().__init__()
        if activation is not nn.GELU or iou_head_depth != 3 or iou_head_hidden_dim != transformer_dim:
            raise NotImplementedError("the B200 decoder implements GELU up-scaling and a 3-layer IoU head of width transformer_dim")
        self.transformer_dim = transformer_dim
        self.transformer = transformer
        self.num_multimask_outputs = num_multimask_outputs
        self.iou_token = nn.Embedding(1, transformer_dim)
        self.num_mask_tokens = num_multimask_outputs + 1
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, transformer_dim)
        self.output_upscaling = nn.Sequential(
            nn.ConvTranspose2d(transformer_dim, transformer_dim // 4, kernel_size=2, stride=2),
            LayerNorm2d(transformer_dim // 4),
            activation(),
            nn.ConvTranspose2d(transformer_dim // 4, transformer_dim // 8, kernel_size=2, stride=2),
            activation(),
        )
        self.output_hypernetworks_mlps = nn.ModuleList(
            MLP(transformer_dim, transformer_dim, transformer_dim // 8, 3) for _ in range(self.num_mask_tokens))
        self.iou_prediction_head = MLP(transformer_dim, iou_head_hidden_dim, self.num_mask_tokens, iou_head_depth)
        self.edge_token = nn.Embedding(1, transformer_dim)
        self.edge_mlp = MLP(transformer_dim, transformer_dim, transformer_dim // 8, 3)
        self.num_mask_tokens = self.num_mask_tokens + 1
        self.embedding_encoder = nn.Sequential(
            nn.ConvTranspose2d(transformer_dim, transformer_dim // 4, 2, 2),
            LayerNorm2d(transformer_dim // 4),
            nn.GELU(),
            nn.ConvTranspose2d(transformer_dim // 4, transformer_dim // 8, 2, 2),
        )
        self.embedding_maskfeature = nn.Sequential(
            nn.ConvTranspose2d(transformer_dim // 8, transformer_dim // 4, 3, 1, 1),
            LayerNorm2d(transformer_dim // 4),
            nn.GELU(),
            nn.ConvTranspose2d(transformer_dim // 4, transformer_dim // 8, 3, 1, 1),
        )
        self.sigmoid = nn.Sigmoid()
        self.max_chunk = 32             # images per C-ABI call; bounds the workspace (~88 MB / image at g = 64)
        self._handles: Dict[Tuple[int, int], int] = {}
        self._packed: Optional[Dict[str, torch.Tensor]] = None
        self._packed_key = None
        self._registered = set()
        self._workspace: Optional[torch.Tensor] = None
        self.last_launch_count = 0
        self.last_hs: Optional[torch.Tensor] = None

    _RUNTIME_FIELDS = ("_packed", "_packed_key", "_workspace", "last_hs")

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in self._RUNTIME_FIELDS:
            state[k] = None
        state["_handles"] = {}
        state["_registered"] = set()
        return state

    def __del__(self):
        try:
            for h in self._handles.values():
                _lib.load().cvlm_decoder_destroy(h)
        except Exception:
            pass

    def _depth(self) -> int:
        return len(self.transformer.layers)

    def _weights_key(self, device):
        return (str(device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _handle_for(self, grid: int, cond_tokens: int, device) -> int:
        lib = _lib.load()
        key = self._weights_key(device)
        if self._packed is None or self._packed_key != key:
            self._packed = pack_decoder_weights(dict(self.state_dict()), self._depth(), device)
            self._packed_key = key
            self._registered = set()
        hk = (grid, cond_tokens)
        if hk not in self._handles:
            cc = _lib.DecoderConfigC()
            cc.dim, cc.heads = self.transformer_dim, int(getattr(self.transformer, "num_heads", 8))
            cc.mlp_dim = int(self.transformer.layers[0].mlp.lin1.out_features)
            cc.depth, cc.num_mask_tokens = self._depth(), self.num_mask_tokens - 1
            cc.grid, cc.cond_tokens = grid, cond_tokens
            cc.ln_eps = float(self.transformer.norm_final_attn.eps)
            cc.ln2d_eps = float(self.output_upscaling[1].eps)
            h = C.c_void_p()
            _lib.check(lib.cvlm_decoder_create(C.byref(cc), C.byref(h)), "cvlm_decoder_create")
            self._handles[hk] = h.value
        if hk not in self._registered:
            for name, t in self._packed.items():
                _lib.check(lib.cvlm_decoder_set_tensor(self._handles[hk], name.encode(), t.data_ptr(),
                                                       t.numel() * t.element_size()), f"decoder set_tensor({name})")
            self._registered.add(hk)
        return self._handles[hk]

    def predict_all(self, image_embeddings: torch.Tensor, image_pe: torch.Tensor, sparse_prompt_embeddings: torch.Tensor,
                    dense_prompt_embeddings: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """predict_masks (mask_decoder_edge.py:140-190) for a batch: masks [B,4,4g,4g], edge [B,1,4g,4g], iou [B,4]."""
        if not image_embeddings.is_cuda:
            raise RuntimeError("MaskDecoder (B200) needs CUDA tensors: there is no CPU path")
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (
                image_embeddings, sparse_prompt_embeddings, dense_prompt_embeddings)):
            # parameters with requires_grad=True (the default of a freshly built module) do not block inference;
            # a differentiable INPUT does, because no gradient would reach it
            raise RuntimeError("MaskDecoder (B200) is inference only: an input requires grad; detach it or call under "
                               "torch.no_grad()")
        dev = image_embeddings.device
        B, Cc, g, g2 = image_embeddings.shape
        if Cc != self.transformer_dim or g != g2:
            raise ValueError(f"image_embeddings must be [B,{self.transformer_dim},g,g], got {tuple(image_embeddings.shape)}")
        if sparse_prompt_embeddings.shape[0] != B or sparse_prompt_embeddings.shape[2] != Cc:
            raise ValueError("sparse_prompt_embeddings must be [B, tokens, transformer_dim] with one prompt per image")
        if tuple(image_pe.shape) != (1, Cc, g, g):
            raise ValueError(f"image_pe must be [1,{Cc},{g},{g}]")
        nc = int(sparse_prompt_embeddings.shape[1])
        emb = image_embeddings.detach().float().contiguous()
        pe = image_pe.detach().float().contiguous()
        sparse = sparse_prompt_embeddings.detach().float().contiguous()
        dense = dense_prompt_embeddings.detach().float()
        if dense.shape[0] not in (1, B) or tuple(dense.shape[1:]) != (Cc, g, g):
            raise ValueError(f"dense_prompt_embeddings must be [1 or B,{Cc},{g},{g}]")
        if dense.shape[0] == B and B > 1 and dense.stride(0) == 0:
            dense = dense[:1]                             # the no_mask_embed broadcast (models/sam_maskdecoder_edge.py:307-309)
        dense = dense.contiguous()
        lib = _lib.load()
        h = self._handle_for(g, nc, dev)
        masks = torch.empty(B, 4, 4 * g, 4 * g, device=dev, dtype=torch.float32)
        edge = torch.empty(B, 1, 4 * g, 4 * g, device=dev, dtype=torch.float32)
        iou = torch.empty(B, 4, device=dev, dtype=torch.float32)
        hs = torch.empty(B, NUM_TOKENS, Cc, device=dev, dtype=torch.float32)
        launches = 0
        for s in range(0, B, self.max_chunk):
            e = min(B, s + self.max_chunk)
            need = _lib.check(lib.cvlm_decoder_workspace_bytes(h, e - s), "cvlm_decoder_workspace_bytes")
            if self._workspace is None or self._workspace.numel() < need or self._workspace.device != dev:
                # a CUDA graph that captured the old buffer keeps its own reference (graphed.py): replacing is safe
                self._workspace = torch.empty(need, dtype=torch.uint8, device=dev)
            ws = self._workspace
            db = dense.shape[0]
            dptr = dense.data_ptr() if db == 1 else dense[s:e].data_ptr()
            with torch.cuda.device(dev):        # the C ABI launches on the CUDA current device
                _lib.check(lib.cvlm_decoder_forward(h, emb[s:e].data_ptr(), pe.data_ptr(), sparse[s:e].data_ptr(), dptr,
                                                    1 if db == 1 else e - s, e - s, masks[s:e].data_ptr(), edge[s:e].data_ptr(),
                                                    iou[s:e].data_ptr(), hs[s:e].data_ptr(), ws.data_ptr(), ws.numel(),
                                                    _lib.stream_ptr(dev)), "cvlm_decoder_forward")
            launches += lib.cvlm_decoder_last_launches(h)
        self.last_launch_count = launches
        self.last_hs = hs
        return masks, edge, iou

    def forward(self, image_embeddings: torch.Tensor, interm_embeddings: List[torch.Tensor], image_pe: torch.Tensor,
                sparse_prompt_embeddings: torch.Tensor, dense_prompt_embeddings: torch.Tensor,
                multimask_output: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        # interm_embeddings is accepted and ignored, like the reference (mask_decoder_edge.py:99, never read)
        masks, edge, iou = self.predict_all(image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings)
        sl = slice(1, None) if multimask_output else slice(0, 1)        # mask_decoder_edge.py:133-139
        return masks[:, sl, :, :], edge, iou[:, sl]


MaskDecoder_Edge = MaskDecoder      # the name the cascade imports it under (models/sam_maskdecoder_edge.py:10)


def build_mask_decoder(seed: Optional[int] = None) -> MaskDecoder:
    """The decoder as the cascade constructs it (models/sam_maskdecoder_edge.py:137-148).  With ``seed`` the parameters are
    re-drawn from a seeded normal (std 0.05, norm weights around 1): synthetic weights for benchmarks."""
    dec = MaskDecoder(num_multimask_outputs=3, transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256,
                      transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8)).eval()
    if seed is not None:
        gen = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for n, p in dec.named_parameters():
                p.copy_(torch.randn(p.shape, generator=gen) * 0.05 + (1.0 if (p.dim() == 1 and n.endswith("weight")) else 0.0))
    return dec


def bilinear_resize(x: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """F.interpolate(x, size, mode="bilinear", align_corners=False) for fp32 NCHW CUDA tensors (csrc/decoder.cu)."""
    if not x.is_cuda:
        raise RuntimeError("bilinear_resize needs a CUDA tensor: there is no CPU path")
    x = x.float().contiguous()
    n, c, h, w = x.shape
    out = torch.empty(n, c, size[0], size[1], device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().cvlm_op_bilinear_resize(x.data_ptr(), out.data_ptr(), n * c, h, w, size[0], size[1],
                                                       _lib.stream_ptr(x.device)), "cvlm_op_bilinear_resize")
    return out


def postprocess_masks(masks: torch.Tensor, img_size: int, input_size: int, original_size) -> torch.Tensor:
    """models/sam_maskdecoder_edge.py:359-388: up-sample to the encoder's input size, crop the padding, resize to the
    original image size."""
    if isinstance(original_size, int):
        original_size = (original_size, original_size)
    masks = bilinear_resize(masks, (img_size, img_size))
    if input_size == img_size and tuple(original_size) == (img_size, img_size):
        return masks        # the second interpolation is the identity at equal sizes (scale 1, zero offset)
    masks = masks[..., :input_size, :input_size]
    return bilinear_resize(masks, tuple(original_size))
