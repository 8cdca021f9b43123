"""Gradient goldens for the adapter-only training path, container only.

    python oracle/make_golden_grads.py [--full]      # --full adds ViT-H at 1024^2 (a few minutes, ~25 GB of RAM)

Runs the UNMODIFIED reference ``ImageEncoderViT`` under autograd with the reference training script's freeze rule
(train_ovcos_maskdecoder_edge.py:301-303: every ``image_encoder`` parameter without "prompt_generator" in its name is
frozen), loss = sum(out * c_out) + sum_k sum(interm_k * c_k) with seeded cotangents (oracle.encoder_oracle.cotangents),
and writes the gradient of every ``prompt_generator.*`` parameter to tests/golden/encoder_grads_<name>.npz.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from camouflaged_vlm_b200.config import PRESETS  # noqa: E402
from camouflaged_vlm_b200.synth import seeded_images, seeded_state_dict  # noqa: E402
from oracle.encoder_oracle import cotangents  # noqa: E402
from oracle.ref_import import build_reference_encoder  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def reference_grads(name, batch, seed=0, with_interm=True):
    cfg = PRESETS[name]
    ref = build_reference_encoder(cfg).train()          # no dropout / batch-norm: train() == eval() numerically
    ref.load_state_dict(seeded_state_dict(cfg, seed), strict=True)
    for n, p in ref.named_parameters():
        # the script's filter sees names like "image_encoder.blocks.0..."; inside the encoder that is:
        p.requires_grad_("prompt_generator" in n)
    x = seeded_images(cfg, batch, seed)
    c_out, c_int = cotangents(cfg, batch, seed)
    t0 = time.perf_counter()
    out, interm = ref(x, interm=True)
    loss = (out * c_out).sum()
    if with_interm:
        for t, c in zip(interm, c_int):
            loss = loss + (t * c).sum()
    loss.backward()
    grads = {n: p.grad.detach().numpy() for n, p in ref.named_parameters() if p.requires_grad}
    print(f"[golden grads] {name} B={batch}: {time.perf_counter() - t0:.1f} s, {len(grads)} tensors, loss {float(loss):.4f}",
          flush=True)
    return grads, float(loss)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    args = ap.parse_args()
    cases = [("tiny", 2), ("small_hd80", 1), ("small_hd64", 1)]
    if args.full:
        cases.append(("vit_h", 1))
    for name, batch in cases:
        grads, loss = reference_grads(name, batch)
        np.savez_compressed(os.path.join(GOLDEN, f"encoder_grads_{name}.npz"), loss=np.float64(loss), **grads)


if __name__ == "__main__":
    main()
