#!/usr/bin/env python
"""Per-kernel-kind time of one adapter-training step (forward_train + backward) of the ViT-H encoder, CUDA events around
every launch (cvlm_encoder_set_profiling).   python tools/train_bench.py [--batch 8] [--model vit_h]"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from camouflaged_vlm_b200 import PRESETS, _lib, build_encoder
from camouflaged_vlm_b200.synth import seeded_state_dict


def profile(model):
    lib = _lib.load()
    ms = (C.c_float * len(_lib.PROF_KINDS))()
    n = (C.c_int * len(_lib.PROF_KINDS))()
    _lib.check(lib.cvlm_encoder_get_profile(model._handle, ms, n))
    return {k: {"ms": round(ms[i], 3), "launches": n[i]} for i, k in enumerate(_lib.PROF_KINDS)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--model", default="vit_h")
    args = ap.parse_args()
    cfg = PRESETS[args.model]
    m = build_encoder(cfg)
    m.load_state_dict(seeded_state_dict(cfg, 0), strict=True)
    m = m.cuda().train()
    for n, p in m.named_parameters():
        p.requires_grad_("prompt_generator" in n)
    x = torch.randn(args.batch, 3, cfg.img_size, cfg.img_size, device="cuda")
    cot = torch.randn(args.batch, cfg.out_chans, cfg.grid, cfg.grid, device="cuda")
    for _ in range(2):
        m.zero_grad()
        (m(x) * cot).sum().backward()
    torch.cuda.synchronize()
    lib = _lib.load()
    lib.cvlm_encoder_set_profiling(m._handle, 1)
    m.zero_grad()
    out = m(x)
    torch.cuda.synchronize()
    fwd = profile(m)
    (out * cot).sum().backward()
    torch.cuda.synchronize()
    bwd = profile(m)
    lib.cvlm_encoder_set_profiling(m._handle, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m.zero_grad()
        (m(x) * cot).sum().backward()
    e1.record()
    torch.cuda.synchronize()
    rep = {"batch": args.batch, "model": args.model, "step_ms": e0.elapsed_time(e1) / 3,
           "forward_train_ms_by_kind": fwd, "backward_ms_by_kind": bwd,
           "forward_train_ms": sum(v["ms"] for v in fwd.values()), "backward_ms": sum(v["ms"] for v in bwd.values())}
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
