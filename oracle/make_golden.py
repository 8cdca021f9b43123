"""Generate tests/golden/*.npz by running the UNMODIFIED reference encoder (container only).

    python oracle/make_golden.py [--full]       # --full also runs ViT-H at 1024^2 (about 30 s + 3 GB)

Weights and inputs come from ``camouflaged_vlm_b200.synth`` (numpy RandomState seeded per tensor
name), so the fixtures only need to hold OUTPUTS: the tests regenerate the same weights / images
and compare.  The ViT-H fixture stores a strided sub-sample of the output and of the four
intermediate embeddings to stay small.
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
from oracle.ref_import import build_reference_encoder  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
VITH_STRIDE = 8       # keep out[:, :, ::8, ::8]
VITH_CH_STRIDE = 16   # keep interm[:, ::16, ::8, ::8]


def run_reference(name: str, batch: int, seed: int = 0):
    cfg = PRESETS[name]
    ref = build_reference_encoder(cfg).eval()
    missing = ref.load_state_dict(seeded_state_dict(cfg, seed), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    x = seeded_images(cfg, batch, seed)
    with torch.no_grad():
        t0 = time.perf_counter()
        out, interm = ref(x, interm=True)
        dt = time.perf_counter() - t0
    print(f"[golden] {name} B={batch}: {dt:.2f} s, out {tuple(out.shape)}, std {out.std():.4f}")
    return out, interm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    for name, batch in (("tiny", 2), ("small_hd80", 1), ("small_hd64", 1)):
        out, interm = run_reference(name, batch)
        np.savez_compressed(os.path.join(GOLDEN, f"encoder_{name}.npz"),
                            out=out.numpy(),
                            **{f"interm{i}": t.contiguous().numpy() for i, t in enumerate(interm)})
    # downstream probe: the reference MaskDecoder_Edge on a seeded random embedding (pins the decoder oracle)
    from oracle import mask_decoder_oracle as mdo
    from oracle.ref_import import load_reference_decoder_modules
    md, tr = load_reference_decoder_modules()
    dec = md.MaskDecoder(num_multimask_outputs=3, transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256,
                         transformer=tr.TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8)).eval()
    dsd = mdo.seeded_decoder_state(0)
    dec.load_state_dict({k: v for k, v in dsd.items() if not k.startswith("probe.")}, strict=True)

    def ref_decode(emb):
        g = emb.shape[-1]
        dense = dsd["probe.no_mask_embed"].reshape(1, -1, 1, 1).expand(1, -1, g, g)
        with torch.no_grad():
            masks, _, _ = dec(image_embeddings=emb, interm_embeddings=[], image_pe=mdo.dense_pe(dsd, g),
                              sparse_prompt_embeddings=dsd["probe.sparse"], dense_prompt_embeddings=dense,
                              multimask_output=False)
        return masks

    emb = torch.randn(1, 256, 16, 16, generator=torch.Generator().manual_seed(0))
    # every output of the reference's predict_masks (all four mask tokens, edge map, IoU head) for the decoder parity tests
    dense = dsd["probe.no_mask_embed"].reshape(1, -1, 1, 1).expand(1, -1, 16, 16)
    with torch.no_grad():
        m_all, e_all, iou = dec.predict_masks(image_embeddings=emb, edge_embeddings=dec.embedding_encoder(emb),
                                              image_pe=mdo.dense_pe(dsd, 16), sparse_prompt_embeddings=dsd["probe.sparse"],
                                              dense_prompt_embeddings=dense)
    np.savez_compressed(os.path.join(GOLDEN, "decoder_small.npz"), emb=emb.numpy(), logits=ref_decode(emb).numpy(),
                        masks_all=m_all.numpy(), edge=e_all.numpy(), iou=iou.numpy())
    if args.full:
        out, interm = run_reference("vit_h", 1)
        # full reference cascade tail on the reference embedding: low-res mask logits [1,1,256,256]
        np.savez_compressed(os.path.join(GOLDEN, "decoder_vit_h_logits.npz"),
                            logits=ref_decode(out).numpy().astype(np.float16))
        s, c = VITH_STRIDE, VITH_CH_STRIDE
        np.savez_compressed(os.path.join(GOLDEN, "encoder_vit_h_sub.npz"),
                            out=out[:, :, ::s, ::s].contiguous().numpy(),
                            out_norm=np.float64(out.double().norm().item()),
                            out_mean=np.float64(out.double().mean().item()),
                            **{f"interm{i}": t[:, ::c, ::s, ::s].contiguous().numpy()
                               for i, t in enumerate(interm)})


if __name__ == "__main__":
    main()
