"""Drop-in replacement for the reference ``ImageEncoderViT``
(models/mmseg/models/sam/image_encoder.py:25-155): same constructor, same parameter tree / state_dict
keys, same ``forward(x, interm=False)`` contract, same public attributes -- but the forward runs the
hand-written sm_100a kernels of libcvlm_b200.so through the C ABI in include/cvlm.h.

The sub-modules below (PatchEmbed, Block, Attention, PromptGenerator, ...) are parameter containers
that reproduce the reference's module tree so that ``state_dict()`` / ``load_state_dict(strict=True)``
and name filters such as ``"prompt_generator" in name`` (train_ovcos_maskdecoder_edge.py:301-303) keep
working.  They carry no compute: the product path never calls them.

There is no CPU path and no PyTorch fallback: ``forward`` raises if the input is not on a CUDA
device or if the library has not been built.  With grad mode on and only ``prompt_generator.*`` parameters
requiring grad (the reference's training configuration, train_ovcos_maskdecoder_edge.py:301-303) ``forward`` is
differentiable with respect to those parameters (camouflaged_vlm_b200/backward.py).  ``forward_check`` is the explicit fp32 torch check
path (check_path.py), for parity testing only.
"""
import ctypes as C
import math
from functools import partial
from typing import Any, Dict, List, Optional, Tuple, Type

import torch
import torch.nn as nn

from . import _lib
from .check_path import encoder_forward_check
from .common import LayerNorm2d, MLPBlock
from .config import EncoderConfig
from .packing import pack_weights, rank_pad


class PatchEmbed(nn.Module):
    """Parameter container for image_encoder.py:628-659 (Conv2d k=stride=patch)."""

    def __init__(self, kernel_size=(16, 16), stride=(16, 16), padding=(0, 0), in_chans=3, embed_dim=768):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride, padding=padding)


class PatchEmbed2(nn.Module):
    """Parameter container for image_encoder.py:355-380."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class Attention(nn.Module):
    """Parameter container for image_encoder.py:449-486."""

    def __init__(self, dim, num_heads=8, qkv_bias=True, use_rel_pos=False, rel_pos_zero_init=True, input_size=None):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.use_rel_pos = use_rel_pos
        if use_rel_pos:
            assert input_size is not None, "Input size must be provided if using relative positional encoding."
            self.rel_pos_h = nn.Parameter(torch.zeros(2 * input_size[0] - 1, head_dim))
            self.rel_pos_w = nn.Parameter(torch.zeros(2 * input_size[1] - 1, head_dim))


class Block(nn.Module):
    """Parameter container for image_encoder.py:383-428; ``window_size`` is read by callers (:147)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=True, norm_layer=nn.LayerNorm, act_layer=nn.GELU,
                 use_rel_pos=False, rel_pos_zero_init=True, window_size=0, input_size=None):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, use_rel_pos=use_rel_pos,
                              rel_pos_zero_init=rel_pos_zero_init,
                              input_size=input_size if window_size == 0 else (window_size, window_size))
        self.norm2 = norm_layer(dim)
        self.mlp = MLPBlock(embedding_dim=dim, mlp_dim=int(dim * mlp_ratio), act=act_layer)
        self.window_size = window_size


def _trunc_normal_(t: torch.Tensor, std: float) -> None:
    """Same distribution as the reference's trunc_normal_(std=.02) (:164-215): N(0, std) cut at +-2."""
    nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)


class PromptGenerator(nn.Module):
    """Parameter container for the EVP adapter, image_encoder.py:218-276 (the reference's own
    ``forward`` :298-330 is dead code and is not reproduced)."""

    def __init__(self, scale_factor, prompt_type, embed_dim, tuning_stage, depth, input_type, freq_nums,
                 handcrafted_tune, embedding_tune, adaptor, img_size, patch_size):
        super().__init__()
        self.scale_factor, self.prompt_type, self.embed_dim = scale_factor, prompt_type, embed_dim
        self.input_type, self.freq_nums, self.tuning_stage, self.depth = input_type, freq_nums, tuning_stage, depth
        self.handcrafted_tune, self.embedding_tune, self.adaptor = handcrafted_tune, embedding_tune, adaptor
        r = embed_dim // scale_factor
        self.shared_mlp = nn.Linear(r, embed_dim)
        self.embedding_generator = nn.Linear(embed_dim, r)
        for i in range(depth):
            setattr(self, f"lightweight_mlp_{i}", nn.Sequential(nn.Linear(r, r), nn.GELU()))
        self.prompt_generator = PatchEmbed2(img_size=img_size, patch_size=patch_size, in_chans=3, embed_dim=r)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):   # image_encoder.py:263-276
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)
        elif isinstance(m, nn.Conv2d):
            fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
            m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
            if m.bias is not None:
                m.bias.data.zero_()


class ImageEncoderViT(nn.Module):
    def __init__(
        self,
        img_size: int = 1024,
        patch_size: int = 16,
        in_chans: int = 3,
        embed_dim: int = 768,
        depth: int = 12,
        num_heads: int = 12,
        mlp_ratio: float = 4.0,
        out_chans: int = 256,
        qkv_bias: bool = True,
        norm_layer: Type[nn.Module] = nn.LayerNorm,
        act_layer: Type[nn.Module] = nn.GELU,
        use_abs_pos: bool = True,
        use_rel_pos: bool = False,
        rel_pos_zero_init: bool = True,
        window_size: int = 0,
        global_attn_indexes: Tuple[int, ...] = (),
    ) -> None:
        super().__init__()
        probe = norm_layer(embed_dim)
        if not isinstance(probe, nn.LayerNorm) or not probe.elementwise_affine:
            raise NotImplementedError("ImageEncoderViT (B200): norm_layer must build an affine nn.LayerNorm")
        act = act_layer()
        if not isinstance(act, nn.GELU) or getattr(act, "approximate", "none") != "none":
            raise NotImplementedError("ImageEncoderViT (B200): act_layer must be exact-erf nn.GELU")
        if in_chans != 3:
            raise NotImplementedError("the adapter's PatchEmbed2 hard-codes 3 input channels (image_encoder.py:258)")
        self.img_size = img_size
        self.embed_dim = embed_dim
        self.depth = depth
        self._ln_eps = float(probe.eps)
        self._cfg = EncoderConfig(
            img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim, depth=depth,
            num_heads=num_heads, mlp_ratio=mlp_ratio, out_chans=out_chans, qkv_bias=qkv_bias,
            use_abs_pos=use_abs_pos, use_rel_pos=use_rel_pos, window_size=window_size,
            global_attn_indexes=tuple(int(i) for i in global_attn_indexes))

        self.patch_embed = PatchEmbed(kernel_size=(patch_size, patch_size), stride=(patch_size, patch_size),
                                      in_chans=in_chans, embed_dim=embed_dim)
        self.pos_embed: Optional[nn.Parameter] = None
        if use_abs_pos:
            self.pos_embed = nn.Parameter(torch.zeros(1, img_size // patch_size, img_size // patch_size, embed_dim))
        self.blocks = nn.ModuleList()
        for i in range(depth):
            self.blocks.append(Block(
                dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, norm_layer=norm_layer,
                act_layer=act_layer, use_rel_pos=use_rel_pos, rel_pos_zero_init=rel_pos_zero_init,
                window_size=window_size if i not in global_attn_indexes else 0,
                input_size=(img_size // patch_size, img_size // patch_size)))
        self.neck = nn.Sequential(
            nn.Conv2d(embed_dim, out_chans, kernel_size=1, bias=False),
            LayerNorm2d(out_chans),
            nn.Conv2d(out_chans, out_chans, kernel_size=3, padding=1, bias=False),
            LayerNorm2d(out_chans),
        )
        # hard-coded by the reference, image_encoder.py:116-123
        self.scale_factor = 32
        self.prompt_type = 'highpass'
        self.high_pass = "gemm"      # or "fft": PromptGenerator.fft through torch.fft in fp32 (see _run)
        self.tuning_stage = 1234
        self.input_type = 'fft'
        self.freq_nums = 0.25
        self.handcrafted_tune = True
        self.embedding_tune = True
        self.adaptor = 'adaptor'
        self.prompt_generator = PromptGenerator(self.scale_factor, self.prompt_type, self.embed_dim,
                                                self.tuning_stage, self.depth, self.input_type, self.freq_nums,
                                                self.handcrafted_tune, self.embedding_tune, self.adaptor,
                                                img_size, patch_size)
        self.num_stages = self.depth
        self.out_indices = tuple(range(self.num_stages))

        # runtime state (not part of state_dict)
        self.max_chunk = 32            # images per C-ABI call; bounds the workspace (~85 MB / image for ViT-H)
        self._handle: Optional[int] = None
        self._packed: Optional[Dict[str, torch.Tensor]] = None
        self._packed_key = None
        self._workspace: Optional[torch.Tensor] = None
        self._stage = None             # forward_host: two device staging buffers + a copy stream
        self._copy_stream = None
        self._packed_bwd = None        # adapter training: transposed weight copies, backward workspace
        self._packed_bwd_key = None
        self._workspace_bwd = None
        self.last_launch_count = 0
        self.last_backward_launch_count = 0

    # ------------------------------------------------------------------ packed-weight cache
    @property
    def config(self) -> EncoderConfig:
        return self._cfg

    def _weights_key(self, device):
        return (str(device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _ensure_ready(self, device: torch.device) -> None:
        key = self._weights_key(device)
        lib = _lib.load()
        if self._handle is None:
            cfg = self._cfg
            cc = _lib.EncoderConfigC()
            cc.img_size, cc.patch_size, cc.in_chans = cfg.img_size, cfg.patch_size, cfg.in_chans
            cc.embed_dim, cc.depth, cc.num_heads = cfg.embed_dim, cfg.depth, cfg.num_heads
            cc.mlp_dim, cc.out_chans, cc.window_size = cfg.mlp_dim, cfg.out_chans, cfg.window_size
            for i in range(min(cfg.depth, _lib.MAX_DEPTH)):
                cc.is_global[i] = 1 if cfg.block_window(i) == 0 else 0
            cc.use_abs_pos = 1 if cfg.use_abs_pos else 0
            cc.rank, cc.rank_pad = cfg.rank, rank_pad(cfg.rank)
            cc.ln_eps, cc.ln2d_eps = self._ln_eps, float(self.neck[1].eps)
            h = C.c_void_p()
            _lib.check(lib.cvlm_encoder_create(C.byref(cc), C.byref(h)), "cvlm_encoder_create")
            self._handle = h.value
        if self._packed is None or self._packed_key != key:
            sd = {k: v for k, v in self.state_dict().items()}
            self._packed = pack_weights(sd, self._cfg, device, dtype=torch.bfloat16)
            for name, t in self._packed.items():
                _lib.check(lib.cvlm_encoder_set_tensor(self._handle, name.encode(), t.data_ptr(),
                                                       t.numel() * t.element_size()), f"set_tensor({name})")
            self._packed_key = key

    _RUNTIME_FIELDS = ("_handle", "_packed", "_packed_key", "_workspace", "_stage", "_copy_stream", "_packed_bwd",
                       "_packed_bwd_key", "_workspace_bwd")

    def __getstate__(self):
        # copy.deepcopy / pickle must not share the native handle or the device caches
        state = dict(self.__dict__)
        for k in self._RUNTIME_FIELDS:
            state[k] = None
        return state

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().cvlm_encoder_destroy(self._handle)
        except Exception:
            pass

    def _get_workspace(self, batch: int, device) -> torch.Tensor:
        need = _lib.load().cvlm_encoder_workspace_bytes(self._handle, batch)
        _lib.check(need, "cvlm_encoder_workspace_bytes")
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != device:
            self._workspace = None
            self._workspace = torch.empty(need, dtype=torch.uint8, device=device)
        return self._workspace

    # ------------------------------------------------------------------ product path
    def forward(self, x: torch.Tensor, interm: bool = False) -> Tuple[Any, List[Any]]:
        cfg = self._cfg
        if not x.is_cuda:
            raise RuntimeError("ImageEncoderViT (B200): the encoder only runs on a CUDA device; "
                               "there is no CPU fallback (use forward_check for the fp32 torch check path)")
        if x.dim() != 4 or x.shape[1] != cfg.in_chans or x.shape[2] != cfg.img_size or x.shape[3] != cfg.img_size:
            raise ValueError(f"expected input [B, {cfg.in_chans}, {cfg.img_size}, {cfg.img_size}], got {tuple(x.shape)}")
        if torch.is_grad_enabled():
            if x.requires_grad:
                raise RuntimeError("ImageEncoderViT (B200): no gradient with respect to the input image is implemented "
                                   "(detach the input; only prompt_generator.* parameters are trainable)")
            trainable = [n for n, p in self.named_parameters() if p.requires_grad]
            trunk = [n for n in trainable if "prompt_generator" not in n]
            if trainable and not trunk:
                return self._forward_trainable(x, interm)      # adapter-only training
            if trunk and self.training:
                raise RuntimeError(
                    "ImageEncoderViT (B200) trains the adapter only: freeze the trunk the way the reference's training "
                    "script does (train_ovcos_maskdecoder_edge.py:301-303: requires_grad_(False) for every encoder "
                    f"parameter without 'prompt_generator' in its name; {len(trunk)} such parameters still require grad, "
                    f"e.g. {trunk[0]}), or call under torch.no_grad() / .eval() for inference")
            if trunk:
                # eval() without no_grad(): the reference would build a graph nobody uses; here the result is detached
                # (Python's default warning filter shows this once per call site)
                import warnings
                warnings.warn("ImageEncoderViT (B200): parameters of the frozen trunk still have requires_grad=True; "
                              "running inference, the output is detached (no gradient will reach this module)",
                              stacklevel=2)
        with torch.no_grad():
            x = x.detach().to(torch.float32).contiguous()
            B, g, D = x.shape[0], cfg.grid, cfg.embed_dim
            out = torch.empty(B, cfg.out_chans, g, g, device=x.device, dtype=torch.float32)
            n_glob = sum(1 for i in range(cfg.depth) if cfg.block_window(i) == 0)
            interms = [torch.empty(B, g, g, D, device=x.device, dtype=torch.float32) for _ in range(n_glob)] if interm else []
            self._run(x, out, interms)
        if interm:
            return out, [t.permute(0, 3, 1, 2) for t in interms]     # views, like image_encoder.py:148
        return out

    def _forward_trainable(self, x: torch.Tensor, interm: bool):
        """Differentiable forward: gradients for the ``prompt_generator.*`` parameters (camouflaged_vlm_b200/backward.py)."""
        from .backward import EncoderTrainFunction, adapter_parameter_names
        params = dict(self.named_parameters())
        res = EncoderTrainFunction.apply(self, x, *[params[n] for n in adapter_parameter_names(self._cfg)])
        out, interms = res[0], res[1:]
        if interm:
            return out, [t.permute(0, 3, 1, 2) for t in interms]     # views, like image_encoder.py:148
        return out

    def _ensure_backward_ready(self, device: torch.device) -> None:
        """Registers the transposed weight copies the backward pass needs (1.3 GB for ViT-H), once per packing."""
        if self._packed_bwd_key == self._packed_key and self._packed_bwd is not None:
            return
        from .backward import pack_backward_weights
        lib = _lib.load()
        self._packed_bwd = pack_backward_weights(dict(self.state_dict()), self._cfg, device, self._packed)
        for name, t in self._packed_bwd.items():
            _lib.check(lib.cvlm_encoder_set_tensor(self._handle, name.encode(), t.data_ptr(),
                                                   t.numel() * t.element_size()), f"set_tensor({name})")
        self._packed_bwd_key = self._packed_key

    def _get_backward_workspace(self, need: int, device) -> torch.Tensor:
        ws = self._workspace_bwd
        if ws is None or ws.numel() < need or ws.device != device:
            self._workspace_bwd = None
            self._workspace_bwd = torch.empty(need, dtype=torch.uint8, device=device)
        return self._workspace_bwd

    def _run(self, x: torch.Tensor, out: torch.Tensor, interms: List[torch.Tensor]) -> None:
        """One or more C-ABI calls (chunks of ``max_chunk`` images) writing into caller-provided device tensors:
        ``out`` [B,out,g,g] and, when not empty, ``interms`` (one [B,g,g,D] NHWC tensor per global block)."""
        self._ensure_ready(x.device)
        lib = _lib.load()
        B = x.shape[0]
        stream = _lib.stream_ptr(x.device)          # torch's current stream on the TENSORS' device
        launches = 0
        for s in range(0, B, self.max_chunk):
            e = min(B, s + self.max_chunk)
            xc = x[s:e]
            ws = self._get_workspace(e - s, x.device)
            iptr = None
            if interms:
                arr = (C.c_void_p * len(interms))(*[t[s:e].data_ptr() for t in interms])
                iptr = C.cast(arr, C.POINTER(C.c_void_p))
            # highpass = NULL: PromptGenerator.fft runs inside the library (csrc/highpass.cu: two bf16 GEMMs with the
            # low-pass projector, absolute error ~1.6e-3 of the image's RMS).  high_pass = "fft" hands the library an fp32
            # torch.fft result instead: for images whose content above a quarter of the sampling rate is weaker than
            # that (smooth / defocused photographs), see DESIGN.md section 5.2.
            hp = None
            if self.high_pass == "fft":
                from .check_path import high_pass as _hp_fft
                with torch.cuda.device(x.device):
                    hp = _hp_fft(xc, self._cfg.freq_nums).contiguous()
            elif self.high_pass != "gemm":
                raise ValueError(f"high_pass must be 'gemm' or 'fft', not {self.high_pass!r}")
            with torch.cuda.device(x.device):       # the C ABI launches on the CUDA current device
                _lib.check(lib.cvlm_encoder_forward(self._handle, xc.data_ptr(), None if hp is None else hp.data_ptr(),
                                                    e - s, out[s:e].data_ptr(), iptr, ws.data_ptr(), ws.numel(), stream),
                           "cvlm_encoder_forward")
            launches += lib.cvlm_encoder_last_launch_count(self._handle)
        self.last_launch_count = launches

    @torch.no_grad()
    def forward_host(self, x_host: torch.Tensor, out_host: Optional[torch.Tensor] = None, chunk: int = 16,
                     interm: bool = False, return_device: bool = False, on_chunk=None):
        """Embeddings for a batch that lives in HOST memory (``x_host`` fp32 [B,3,S,S], ideally pinned): the batch is
        cut into chunks of ``chunk`` images; the host->device copy of chunk i+1 and the device->host copy of the
        embeddings of chunk i-1 run on a copy stream underneath the encoder of chunk i (two device staging buffers).
        Returns ``out_host`` ([B,out,g,g] fp32, pinned if allocated here), valid once the current stream has been
        synchronised.  ``return_device=True`` appends the device copy of the embeddings (e.g. to all-gather them
        across ranks, ``dist.gather_embeddings``); ``interm=True`` appends the list of per-global-block embeddings,
        which stay on the device (NCHW views, like ``forward``).  ``on_chunk(s, e, emb[s:e])`` is called on the main
        stream right after the encoder of images [s, e) has been enqueued (e.g. ``dist.ChunkGatherer.submit``: the
        all-gather of that chunk then runs under the next chunk's encoder)."""
        cfg = self._cfg
        if x_host.is_cuda:
            raise ValueError("forward_host takes a host tensor; call the module directly for device tensors")
        B, g = x_host.shape[0], cfg.grid
        dev = next(self.parameters()).device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None or self._copy_stream.device != dev:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs = self._copy_stream
        emb = torch.empty(B, cfg.out_chans, g, g, device=dev, dtype=torch.float32)
        n_glob = sum(1 for i in range(cfg.depth) if cfg.block_window(i) == 0)
        interms = [torch.empty(B, g, g, cfg.embed_dim, device=dev, dtype=torch.float32) for _ in range(n_glob)] if interm else []
        shape = (min(chunk, B),) + tuple(x_host.shape[1:])
        if getattr(self, "_stage", None) is None or tuple(self._stage.shape[1:]) != shape or self._stage.device != dev:
            self._stage = torch.empty((2,) + shape, device=dev, dtype=torch.float32)     # two staging buffers, back to back
        stage = [self._stage[0], self._stage[1]]
        ready = [torch.cuda.Event() for _ in range(2)]        # staging buffer filled
        # staging buffer consumed by the encoder: kept across calls (like the C entry point), so that the copy stream only
        # waits for the reader of the half it is about to overwrite and back-to-back calls overlap
        if getattr(self, "_stage_free", None) is None or self._stage_free[0] is not self._stage:
            self._stage_free = (self._stage, [torch.cuda.Event() for _ in range(2)], [False, False])
        free, free_recorded = self._stage_free[1], self._stage_free[2]
        spans = [(s, min(B, s + chunk)) for s in range(0, B, chunk)]

        def upload(i):
            s, e = spans[i]
            with torch.cuda.stream(cs):
                if free_recorded[i % 2]:
                    cs.wait_event(free[i % 2])
                stage[i % 2][:e - s].copy_(x_host[s:e], non_blocking=True)
                ready[i % 2].record(cs)

        if out_host is None:
            out_host = torch.empty(B, cfg.out_chans, g, g, dtype=torch.float32, pin_memory=True)
        if (on_chunk is None and self.high_pass == "gemm" and x_host.dtype == torch.float32 and x_host.is_contiguous()
                and out_host.is_contiguous()):
            # the whole pipeline in ONE C-ABI call (cvlm_encoder_forward_host): same copies, same stream choreography
            self._ensure_ready(dev)
            lib = _lib.load()
            ws = self._get_workspace(min(chunk, B), dev)
            iptr = None
            if interms:
                arr = (C.c_void_p * len(interms))(*[t.data_ptr() for t in interms])
                iptr = C.cast(arr, C.POINTER(C.c_void_p))
            with torch.cuda.device(dev):
                _lib.check(lib.cvlm_encoder_forward_host(self._handle, x_host.data_ptr(), B, min(chunk, B), out_host.data_ptr(),
                                                         emb.data_ptr(), iptr, stage[0].data_ptr(), ws.data_ptr(), ws.numel(),
                                                         cs.cuda_stream, main.cuda_stream), "cvlm_encoder_forward_host")
            self.last_launch_count = lib.cvlm_encoder_last_launch_count(self._handle)
            ret = [out_host]     # (x_host must stay alive until the stream has run, as with any asynchronous copy)
            if return_device:
                ret.append(emb)
            if interm:
                ret.append([t.permute(0, 3, 1, 2) for t in interms])
            return ret[0] if len(ret) == 1 else tuple(ret)
        upload(0)
        for i, (s, e) in enumerate(spans):
            if i + 1 < len(spans):
                upload(i + 1)
            main.wait_event(ready[i % 2])
            self._run(stage[i % 2][:e - s], emb[s:e], [t[s:e] for t in interms])
            free[i % 2].record(main)
            free_recorded[i % 2] = True
            if on_chunk is not None:
                on_chunk(s, e, emb[s:e])
            if i + 1 < len(spans):
                with torch.cuda.stream(cs):         # this chunk's embeddings go out under the next chunk's encoder
                    cs.wait_event(free[i % 2])
                    out_host[s:e].copy_(emb[s:e], non_blocking=True)
            else:                                   # the last ones on the compute stream: the copy stream stays free
                out_host[s:e].copy_(emb[s:e], non_blocking=True)   # for the next call's first upload
        main.wait_stream(cs)
        ret = [out_host]
        if return_device:
            ret.append(emb)
        if interm:
            ret.append([t.permute(0, 3, 1, 2) for t in interms])
        return ret[0] if len(ret) == 1 else tuple(ret)

    @torch.no_grad()
    def forward_partial(self, x: torch.Tensor, stop_after_block: int) -> torch.Tensor:
        """Residual stream [B*T, D] after block ``stop_after_block`` (-1: after the stem).  Debug aid."""
        cfg = self._cfg
        x = x.detach().to(torch.float32).contiguous()
        self._ensure_ready(x.device)
        lib = _lib.load()
        B = x.shape[0]
        ws = self._get_workspace(B, x.device)
        xo = torch.empty(B * cfg.tokens, cfg.embed_dim, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(lib.cvlm_encoder_forward_partial(self._handle, x.data_ptr(), None, B, stop_after_block,
                                                        xo.data_ptr(), ws.data_ptr(), ws.numel(),
                                                        _lib.stream_ptr(x.device)), "cvlm_encoder_forward_partial")
        return xo

    # ------------------------------------------------------------------ fp32 check path (not the product)
    @torch.no_grad()
    def forward_check(self, x: torch.Tensor, interm: bool = False):
        packed = pack_weights(dict(self.state_dict()), self._cfg, x.device, dtype=torch.float32)
        return encoder_forward_check(packed, self._cfg, x.float(), interm=interm, ln_eps=self._ln_eps)


def build_encoder(cfg: EncoderConfig, ln_eps: float = 1e-6) -> ImageEncoderViT:
    """Construct the way the reference callers do (models/sam_maskdecoder_edge.py:119-135)."""
    return ImageEncoderViT(
        img_size=cfg.img_size, patch_size=cfg.patch_size, in_chans=cfg.in_chans, embed_dim=cfg.embed_dim,
        depth=cfg.depth, num_heads=cfg.num_heads, mlp_ratio=cfg.mlp_ratio, out_chans=cfg.out_chans,
        qkv_bias=cfg.qkv_bias, norm_layer=partial(nn.LayerNorm, eps=ln_eps), act_layer=nn.GELU,
        use_abs_pos=cfg.use_abs_pos, use_rel_pos=cfg.use_rel_pos, rel_pos_zero_init=True,
        window_size=cfg.window_size, global_attn_indexes=tuple(cfg.global_attn_indexes))
