"""SURVEY.md 8(d), config 2 context numbers (not a pytest file): the reference algorithm as PLAIN PYTORCH on the same
GPU -- the oracle's functional restatement run on cuda in fp32 and under torch.autocast(bf16) -- next to the B200
encoder.  Prints img/s for each and the rel-L2 of (autocast bf16) and (ours) against fp32: the autocast number is the
noise floor a bf16 encoder built from library kernels reaches.  Usage: python tests/torch_gpu_baseline.py [batch] [out.json]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from camouflaged_vlm_b200 import PRESETS, build_encoder
from camouflaged_vlm_b200.synth import seeded_images, seeded_state_dict
from oracle import encoder_oracle as orc


def timeit(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    cfg = PRESETS["vit_h"]
    sd = seeded_state_dict(cfg, 0)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    x = seeded_images(cfg, B, 0).cuda()
    rep = {"batch": B}
    with torch.no_grad():
        ref = orc.encoder_forward(sd_gpu, cfg, x, interm=False)
        ms = timeit(lambda: orc.encoder_forward(sd_gpu, cfg, x, interm=False))
        rep["torch_fp32_eager"] = {"ms": ms, "img_per_s": B / ms * 1e3}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ac = orc.encoder_forward(sd_gpu, cfg, x, interm=False)
            ms = timeit(lambda: orc.encoder_forward(sd_gpu, cfg, x, interm=False))
        rep["torch_autocast_bf16"] = {"ms": ms, "img_per_s": B / ms * 1e3, "rel_l2_vs_fp32": orc.rel_l2(ac.float().cpu(), ref.cpu())}
        m = build_encoder(cfg).eval()
        m.load_state_dict(sd, strict=True)
        m = m.cuda()
        ours = m(x)
        ms = timeit(lambda: m(x), iters=10)
        rep["b200_encoder"] = {"ms": ms, "img_per_s": B / ms * 1e3, "rel_l2_vs_fp32": orc.rel_l2(ours.cpu(), ref.cpu())}
    for k, v in rep.items():
        print(k, v)
    if len(sys.argv) > 2:
        json.dump(rep, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
