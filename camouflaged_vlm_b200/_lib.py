"""ctypes binding of libcvlm_b200.so (C ABI declared in include/cvlm.h).

The library is built in-tree by ``build()`` (``make`` in csrc/, nvcc for sm_100a) so that the ``.so``
travels with the repository snapshot.  There is NO fallback: if the library is missing or a call
fails, a RuntimeError is raised.
"""
import ctypes as C
import os
import subprocess
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libcvlm_b200.so")
CSRC = os.path.join(_PKG, "csrc")
MAX_DEPTH = 64
STAT_GROUP_COLS = 128     # include/cvlm.h CVLM_STAT_GROUP_COLS
PROF_KINDS = ("gemm_qkv", "gemm_proj", "gemm_lin1", "gemm_lin2", "gemm_other", "attn_window", "attn_global",
              "layernorm", "other")

_lock = threading.Lock()
_lib = None


class CvlmError(RuntimeError):
    pass


class EncoderConfigC(C.Structure):
    """struct cvlm_encoder_config (include/cvlm.h)."""
    _fields_ = [
        ("img_size", C.c_int), ("patch_size", C.c_int), ("in_chans", C.c_int), ("embed_dim", C.c_int),
        ("depth", C.c_int), ("num_heads", C.c_int), ("mlp_dim", C.c_int), ("out_chans", C.c_int),
        ("window_size", C.c_int), ("is_global", C.c_int * MAX_DEPTH), ("use_abs_pos", C.c_int),
        ("rank", C.c_int), ("rank_pad", C.c_int), ("ln_eps", C.c_float), ("ln2d_eps", C.c_float),
    ]


class DecoderConfigC(C.Structure):
    """struct cvlm_decoder_config (include/cvlm.h)."""
    _fields_ = [("dim", C.c_int), ("heads", C.c_int), ("mlp_dim", C.c_int), ("depth", C.c_int),
                ("num_mask_tokens", C.c_int), ("grid", C.c_int), ("cond_tokens", C.c_int),
                ("ln_eps", C.c_float), ("ln2d_eps", C.c_float)]


class VitConfigC(C.Structure):
    """struct cvlm_vit_config (include/cvlm.h)."""
    _fields_ = [("input_resolution", C.c_int), ("patch_size", C.c_int), ("width", C.c_int), ("layers", C.c_int),
                ("heads", C.c_int), ("output_dim", C.c_int), ("n_ctx", C.c_int), ("ln_eps", C.c_float)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libcvlm_b200.so (no GPU needed)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(_PKG), "include", "cvlm.h"))
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        cmd = ["make", "-C", CSRC, "-j8"] + (["-B"] if force else [])
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            print(res.stdout[-4000:])
            print(res.stderr[-4000:])
        if res.returncode != 0:
            raise CvlmError("building libcvlm_b200.so failed (see output above)")
    return LIB_PATH


_SIGS = {
    "cvlm_last_error": (C.c_char_p, []),
    "cvlm_abi_version": (C.c_int, []),
    "cvlm_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "cvlm_get_option": (C.c_int, [C.c_char_p, C.POINTER(C.c_int)]),
    "cvlm_encoder_create": (C.c_int, [C.POINTER(EncoderConfigC), C.POINTER(C.c_void_p)]),
    "cvlm_encoder_destroy": (C.c_int, [C.c_void_p]),
    "cvlm_encoder_set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]),
    "cvlm_encoder_workspace_bytes": (C.c_longlong, [C.c_void_p, C.c_int]),
    "cvlm_encoder_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                       C.POINTER(C.c_void_p), C.c_void_p, C.c_longlong, C.c_void_p]),
    "cvlm_encoder_forward_partial": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                               C.c_void_p, C.c_longlong, C.c_void_p]),
    "cvlm_encoder_forward_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    "cvlm_encoder_last_launch_count": (C.c_longlong, [C.c_void_p]),
    "cvlm_encoder_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "cvlm_encoder_get_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "cvlm_encoder_saved_bytes": (C.c_longlong, [C.c_void_p, C.c_int]),
    "cvlm_encoder_backward_workspace_bytes": (C.c_longlong, [C.c_void_p, C.c_int]),
    "cvlm_encoder_grad_floats": (C.c_longlong, [C.c_void_p]),
    "cvlm_encoder_grad_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_int)]),
    "cvlm_encoder_forward_train": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_void_p),
                                             C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p]),
    "cvlm_encoder_backward": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_longlong,
                                        C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p]),
    "cvlm_op_attention_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cvlm_op_attention_backward_lse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvlm_op_attention_lse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvlm_op_ln_bwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_longlong, C.c_int, C.c_float, C.c_void_p]),
    "cvlm_op_wgrad_tn_jp": (C.c_int, [C.c_int, C.c_int]),
    "cvlm_op_wgrad_tn_splits": (C.c_int, [C.c_longlong]),
    "cvlm_op_wgrad_tn": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_longlong,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvlm_op_gemm_bf16_dgelu": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p]),
    "cvlm_op_gemm_bf16": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int,
                                    C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int,
                                    C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "cvlm_op_gemm_bf16_ln": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int,
                                       C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int,
                                       C.c_void_p, C.c_longlong, C.c_int, C.c_int,
                                       C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvlm_op_ln_finalize": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_longlong, C.c_void_p]),
    "cvlm_op_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_void_p]),
    "cvlm_op_attention_plain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cvlm_op_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong,
                                    C.c_int, C.c_float, C.c_void_p]),
    "cvlm_op_layernorm2d_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                           C.c_int, C.c_float, C.c_void_p]),
    "cvlm_op_im2col_patch16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cvlm_op_im2col_3x3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cvlm_op_cast_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "cvlm_op_preprocess_u8_scratch_bytes": (C.c_longlong, [C.c_int, C.c_int, C.c_int]),
    "cvlm_op_preprocess_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_float),
                                        C.POINTER(C.c_float), C.c_void_p, C.c_longlong, C.c_void_p]),
    "cvlm_op_high_pass_scratch_bytes": (C.c_longlong, [C.c_int, C.c_int, C.c_int]),
    "cvlm_op_high_pass": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_longlong, C.c_void_p]),
    "cvlm_op_adapter_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cvlm_decoder_create": (C.c_int, [C.POINTER(DecoderConfigC), C.POINTER(C.c_void_p)]),
    "cvlm_decoder_destroy": (C.c_int, [C.c_void_p]),
    "cvlm_decoder_set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]),
    "cvlm_decoder_workspace_bytes": (C.c_longlong, [C.c_void_p, C.c_int]),
    "cvlm_decoder_last_launches": (C.c_longlong, [C.c_void_p]),
    "cvlm_decoder_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "cvlm_vit_create": (C.c_int, [C.POINTER(VitConfigC), C.POINTER(C.c_void_p)]),
    "cvlm_vit_destroy": (C.c_int, [C.c_void_p]),
    "cvlm_vit_set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]),
    "cvlm_vit_workspace_bytes": (C.c_longlong, [C.c_void_p, C.c_int]),
    "cvlm_vit_last_launches": (C.c_longlong, [C.c_void_p]),
    "cvlm_vit_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_longlong, C.c_void_p]),
    "cvlm_op_bilinear_resize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)


def load():
    """dlopen the library and declare every prototype.  Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise CvlmError(
                    f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(there is no CPU or PyTorch fallback for the encoder)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGS.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        msg = load().cvlm_last_error().decode(errors="replace")
        raise CvlmError(f"{what or 'cvlm call'} failed ({rc}): {msg}")
    return rc


def stream_ptr(device=None) -> int:
    """cudaStream_t of torch's current stream ON `device` (the tensors' device, not the current device)."""
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def on_device(t):
    """Context manager: make the tensor's device current for a C-ABI call (function attributes, per-device caches and
    the kernel launches all follow the CUDA current device)."""
    import torch
    return torch.cuda.device(t.device)


def set_option(name: str, value: int) -> None:
    check(load().cvlm_set_option(name.encode(), int(value)), f"cvlm_set_option({name})")


def get_option(name: str) -> int:
    v = C.c_int()
    check(load().cvlm_get_option(name.encode(), C.byref(v)), f"cvlm_get_option({name})")
    return v.value
