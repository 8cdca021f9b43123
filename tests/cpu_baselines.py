"""CPU baselines for the rows built after the encoder: the oracles -- CPU restatements of the reference's
mask decoder (+ postprocess) and Alpha-CLIP / MaPLe visual tower -- timed on this machine's host cores, one image at a time,
at the deployed sizes.  Lives under tests/ because only tests/, smoke() and bench.py may run anything under oracle/.
(The encoder's CPU baseline is bench.py's own `cpu_baseline` / `--impl reference`.)

    python tests/cpu_baselines.py [out.json]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import alpha_clip_oracle as aco
from oracle import mask_decoder_oracle as mdo


def timed(fn, n=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n


def main(out_path=None):
    gen = torch.Generator().manual_seed(0)
    sd = mdo.seeded_decoder_state(0)
    emb = torch.randn(1, 256, 64, 64, generator=gen)
    dt_dec = timed(lambda: mdo.binary_mask(mdo.decode_all(sd, emb)["masks"][:, 0:1], 1024))
    cfg = aco.VitConfig()
    vsd = aco.seeded_vit_state(cfg, 0)
    shared, deep = aco.seeded_prompts(cfg, depth=9)
    img = torch.randn(1, 3, 336, 336, generator=gen)
    alpha = torch.ones(1, 1, 336, 336)
    dt_tow = timed(lambda: aco.vit_maple_forward(vsd, cfg, img, alpha, shared, deep))
    rep = {"cores": torch.get_num_threads(), "kind": "port",
           "mask_decoder": {"value": 1.0 / dt_dec, "unit": "img/s", "sample": "3 images, oracle decode_all + postprocess, 64x64x256 embedding"},
           "alpha_clip_tower": {"value": 1.0 / dt_tow, "unit": "img/s", "sample": "3 images, oracle vit_maple_forward, ViT-L/14@336, one pass"}}
    print(json.dumps(rep))
    if out_path:
        os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
        json.dump(rep, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
