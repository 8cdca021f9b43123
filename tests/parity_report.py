"""Print (and save) the measured parity numbers of the CUDA encoder on this GPU: rel-L2 of the embedding
and the intermediate embeddings against the reference's golden outputs, for every fixture.  GPU only."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from camouflaged_vlm_b200 import PRESETS, build_encoder
from camouflaged_vlm_b200.check_path import high_pass
from camouflaged_vlm_b200.synth import seeded_images, seeded_state_dict
from oracle import encoder_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main(out_path):
    rep = {}
    for name, batch in (("tiny", 2), ("small_hd80", 1), ("small_hd64", 1), ("vit_h", 1)):
        cfg = PRESETS[name]
        sub = name == "vit_h"
        g = np.load(os.path.join(GOLDEN, f"encoder_{name}{'_sub' if sub else ''}.npz"))
        m = build_encoder(cfg).eval()
        m.load_state_dict(seeded_state_dict(cfg), strict=True)
        m = m.cuda()
        x = seeded_images(cfg, batch).cuda()
        with torch.no_grad():
            out, interm = m(x, interm=True)
            chk, chk_i = m.forward_check(x, interm=True) if not sub else (None, None)
        sel = (lambda t: t[:, :, ::8, ::8]) if sub else (lambda t: t)
        seli = (lambda t: t[:, ::16, ::8, ::8]) if sub else (lambda t: t)
        r = {"embedding_rel_l2_bf16_path": orc.rel_l2(sel(out).cpu(), torch.from_numpy(g["out"])),
             "interm_rel_l2_bf16_path": [orc.rel_l2(seli(t).cpu(), torch.from_numpy(g[f"interm{i}"]))
                                         for i, t in enumerate(interm)]}
        if chk is not None:
            r["embedding_rel_l2_fp32_check_path_on_gpu"] = orc.rel_l2(chk.cpu(), torch.from_numpy(g["out"]))
        rep[name] = r
        print(name, json.dumps(r))
        del m
        torch.cuda.empty_cache()
    # downstream probe: mask IoU of the reference decoder (oracle port) on our embedding vs the reference's
    from oracle import mask_decoder_oracle as mdo
    cfg = PRESETS["vit_h"]
    m = build_encoder(cfg).eval()
    m.load_state_dict(seeded_state_dict(cfg), strict=True)
    m = m.cuda()
    with torch.no_grad():
        emb = m(seeded_images(cfg, 1).cuda())
    want = torch.from_numpy(np.load(os.path.join(GOLDEN, "decoder_vit_h_logits.npz"))["logits"].astype(np.float32))
    got = mdo.decode_mask_logits(mdo.seeded_decoder_state(0), emb.cpu())
    rep["vit_h_downstream_mask_iou"] = mdo.iou(mdo.binary_mask(got, 1024), mdo.binary_mask(want, 1024))
    rep["vit_h_downstream_logits_rel_l2"] = orc.rel_l2(got, want)
    print("downstream mask IoU:", rep["vit_h_downstream_mask_iou"], "logits rel-L2:", rep["vit_h_downstream_logits_rel_l2"])
    del m
    torch.cuda.empty_cache()
    # how much does the cuFFT high-pass depend on the batch it is computed in?
    cfg = PRESETS["vit_h"]
    x = seeded_images(cfg, 2).cuda()
    a, b = high_pass(x, 0.25)[1:2], high_pass(x[1:2], 0.25)
    rep["high_pass_batch_dependence_max_abs"] = float((a - b).abs().max())
    print("high_pass batch dependence max abs:", rep["high_pass_batch_dependence_max_abs"])
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    json.dump(rep, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity.json")
