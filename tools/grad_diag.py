#!/usr/bin/env python
"""Per-tensor gradient errors of the adapter-only training path against the oracle's autograd (GPU box diagnostic)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from camouflaged_vlm_b200 import PRESETS, build_encoder
from camouflaged_vlm_b200.synth import seeded_images, seeded_state_dict
from oracle import encoder_oracle as orc


def run(name, batch, with_interm=True):
    cfg = PRESETS[name]
    sd = seeded_state_dict(cfg, 0)
    x = seeded_images(cfg, batch, 0)
    c_out, c_int = orc.cotangents(cfg, batch, 0)
    want = orc.adapter_gradients(sd, cfg, x, c_out, c_int if with_interm else None)
    m = build_encoder(cfg)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    for n, p in m.named_parameters():
        p.requires_grad_("prompt_generator" in n)
    out, interm = m(x.cuda(), interm=True)
    loss = (out * c_out.cuda()).sum()
    if with_interm:
        for t, c in zip(interm, c_int):
            loss = loss + (t * c.cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    errs = {n: orc.rel_l2(p.grad.cpu(), want[n]) for n, p in m.named_parameters() if p.requires_grad}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print(f"{name} B={batch} interm={with_interm}: max {max(errs.values()):.3e}; worst: " +
          ", ".join(f"{k.replace('prompt_generator.', '')} {v:.2e}" for k, v in worst), flush=True)


if __name__ == "__main__":
    for name, batch, wi in (("tiny", 1, True), ("tiny", 2, True), ("tiny", 2, False), ("small_hd64", 1, True), ("small_hd64", 2, True),
                            ("small_hd80", 1, True), ("small_hd80", 3, True)):
        run(name, batch, wi)
