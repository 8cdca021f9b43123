"""Reference outputs for the configuration the headline is quoted on (container only; about 12 minutes of CPU).

    python oracle/make_golden_vit_h_multi.py [--images 16] [--decoder-seeds 3] [--weight-seeds 4]

Runs the UNMODIFIED reference ``ImageEncoderViT`` (ViT-H, 1024^2, oracle/ref_import.py) and the UNMODIFIED
reference ``MaskDecoder`` (mask_decoder_edge.py) and writes ``tests/golden/encoder_vit_h_multi.npz``:

  out_sub[i]        reference embedding of image i of ``seeded_images(vit_h, N, 0)`` (weights seed 0), kept as
                    out[:, ::8, ::8] (fp32) + the full tensor's L2 norm          -> embedding rel-L2 per image
  logits[s, i]      low-res mask logits [256,256] (fp16) of the reference decoder with ``seeded_decoder_state(s)``
                    on the FULL reference embedding of image i                    -> binary-mask IoU per (image, seed)
  wseed_out_sub[k]  reference embedding (sub-sampled the same way) for weights seed k = 1.. on image seed k

The B200 tests re-create weights and images from the same seeds (camouflaged_vlm_b200/synth.py), run the CUDA
encoder at batch 32 in ONE call and compare image by image (tests/test_gpu_parity_b32.py).
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
from oracle import mask_decoder_oracle as mdo  # noqa: E402
from oracle.ref_import import build_reference_encoder, load_reference_decoder_modules  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
STRIDE = 8


def reference_decoder(seed):
    md, tr = load_reference_decoder_modules()
    dec = md.MaskDecoder(num_multimask_outputs=3, transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256,
                         transformer=tr.TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8)).eval()
    dsd = mdo.seeded_decoder_state(seed)
    dec.load_state_dict({k: v for k, v in dsd.items() if not k.startswith("probe.")}, strict=True)
    return dec, dsd


def ref_decode(dec, dsd, emb):
    g = emb.shape[-1]
    dense = dsd["probe.no_mask_embed"].reshape(1, -1, 1, 1).expand(1, -1, g, g)
    with torch.no_grad():
        masks, _, _ = dec(image_embeddings=emb, interm_embeddings=[], image_pe=mdo.dense_pe(dsd, g),
                          sparse_prompt_embeddings=dsd["probe.sparse"], dense_prompt_embeddings=dense,
                          multimask_output=False)
    return masks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--decoder-seeds", type=int, default=3)
    ap.add_argument("--weight-seeds", type=int, default=4, help="weights seeds 1..k-1 get one image each")
    args = ap.parse_args()
    cfg = PRESETS["vit_h"]
    decs = [reference_decoder(s) for s in range(args.decoder_seeds)]
    ref = build_reference_encoder(cfg).eval()
    ref.load_state_dict(seeded_state_dict(cfg, 0), strict=True)
    imgs = seeded_images(cfg, args.images, 0)
    out_sub, out_norm, logits = [], [], [[] for _ in decs]
    for i in range(args.images):
        t0 = time.perf_counter()
        with torch.no_grad():
            out = ref(imgs[i:i + 1])
        out_sub.append(out[0, :, ::STRIDE, ::STRIDE].contiguous().numpy())
        out_norm.append(out.double().norm().item())
        for s, (dec, dsd) in enumerate(decs):
            logits[s].append(ref_decode(dec, dsd, out)[0, 0].numpy().astype(np.float16))
        print(f"[golden] vit_h image {i}: {time.perf_counter() - t0:.1f} s", flush=True)
    w_sub, w_norm = [], []
    for k in range(1, args.weight_seeds):
        ref.load_state_dict(seeded_state_dict(cfg, k), strict=True)
        with torch.no_grad():
            out = ref(seeded_images(cfg, 1, k))
        w_sub.append(out[0, :, ::STRIDE, ::STRIDE].contiguous().numpy())
        w_norm.append(out.double().norm().item())
        print(f"[golden] vit_h weights seed {k} done", flush=True)
    np.savez_compressed(os.path.join(GOLDEN, "encoder_vit_h_multi.npz"),
                        out_sub=np.stack(out_sub), out_norm=np.array(out_norm), logits=np.array(logits),
                        wseed_out_sub=np.stack(w_sub) if w_sub else np.zeros((0,)), wseed_out_norm=np.array(w_norm),
                        stride=np.int64(STRIDE))
    print("[golden] wrote encoder_vit_h_multi.npz:", os.path.getsize(os.path.join(GOLDEN, "encoder_vit_h_multi.npz")), "bytes")


if __name__ == "__main__":
    main()
