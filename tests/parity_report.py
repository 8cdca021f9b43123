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
    # the CUDA mask decoder (8 f1): on the reference's small golden, on seeded full-size inputs against the oracle, and
    # the whole cascade (CUDA encoder -> CUDA decoder) against the reference cascade's logits
    from camouflaged_vlm_b200 import mask_decoder as md
    dsd = mdo.seeded_decoder_state(0)
    dec = md.MaskDecoder(num_multimask_outputs=3, transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256,
                         transformer=md.TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8)).eval()
    dec.load_state_dict({k: v for k, v in dsd.items() if not k.startswith("probe.")}, strict=True)
    dec = dec.cuda()

    def run_dec(e, sparse=None):
        B, _, gg, _ = e.shape
        sp = dsd["probe.sparse"].expand(B, -1, -1) if sparse is None else sparse
        dense = dsd["probe.no_mask_embed"].reshape(1, -1, 1, 1).expand(B, -1, gg, gg)
        with torch.no_grad():
            return dec.predict_all(e.cuda(), mdo.dense_pe(dsd, gg).cuda(), sp.cuda(), dense.cuda())

    gs = np.load(os.path.join(GOLDEN, "decoder_small.npz"))
    mk, ed, io = run_dec(torch.from_numpy(gs["emb"]))
    d = {"small_golden": {"masks_rel_l2": orc.rel_l2(mk.cpu(), torch.from_numpy(gs["masks_all"])),
                          "edge_rel_l2": orc.rel_l2(ed.cpu(), torch.from_numpy(gs["edge"])),
                          "iou_pred_max_abs": float((io.cpu() - torch.from_numpy(gs["iou"])).abs().max()),
                          "mask_iou": mdo.iou(mdo.binary_mask(mk[:, 0:1].cpu(), 256),
                                              mdo.binary_mask(torch.from_numpy(gs["logits"]), 256))}}
    gen = torch.Generator().manual_seed(2)
    e64 = torch.randn(2, 256, 64, 64, generator=gen)
    sp64 = torch.randn(2, 2, 256, generator=gen)
    mk, ed, io = run_dec(e64, sp64)
    full = []
    for b in range(2):
        w = mdo.decode_all(dsd, e64[b:b + 1], sp64[b:b + 1])
        full.append({"masks_rel_l2": orc.rel_l2(mk[b:b + 1].cpu(), w["masks"]), "edge_rel_l2": orc.rel_l2(ed[b:b + 1].cpu(), w["edge"]),
                     "hs_rel_l2": orc.rel_l2(dec.last_hs[b:b + 1].cpu(), w["hs"]),
                     "iou_pred_max_abs": float((io[b:b + 1].cpu() - w["iou"]).abs().max()),
                     "mask_iou": mdo.iou(mdo.binary_mask(mk[b:b + 1, 0:1].cpu(), 1024), mdo.binary_mask(w["masks"][:, 0:1], 1024))})
    d["full_size_vs_oracle"] = full
    mk, _, _ = run_dec(emb.float())
    d["cascade_vit_h_vs_reference"] = {"logits_rel_l2": orc.rel_l2(mk[:, 0:1].cpu(), want),
                                       "mask_iou": mdo.iou(mdo.binary_mask(mk[:, 0:1].cpu(), 1024), mdo.binary_mask(want, 1024))}
    rep["mask_decoder"] = d
    print("mask decoder:", json.dumps(d))
    del m, dec
    torch.cuda.empty_cache()
    # the CUDA Alpha-CLIP / MaPLe tower (8 f2): reference golden (small) and the oracle at ViT-L/14@336
    from camouflaged_vlm_b200 import alpha_clip as ac
    from oracle import alpha_clip_oracle as aco

    def tower(cfg, seed):
        t = ac.VisionTransformer_MaPLe(cfg.input_resolution, cfg.patch_size, cfg.width, cfg.layers, cfg.heads, cfg.output_dim,
                                       design_details={"trainer": "MaPLe", "maple_length": cfg.n_ctx}).eval()
        tsd = aco.seeded_vit_state(cfg, seed)
        t.load_state_dict(tsd, strict=True)
        return t.cuda(), tsd

    gt = np.load(os.path.join(GOLDEN, "alpha_clip_small_hd64.npz"))
    t, tsd = tower(aco.SMALL_HD64, 0)
    sh, dp = aco.seeded_prompts(aco.SMALL_HD64, depth=3)
    with torch.no_grad():
        o = t(torch.from_numpy(gt["image"]).cuda(), torch.from_numpy(gt["alpha"]).cuda(), sh.cuda(), [d.cuda() for d in dp])
    tw = {"small_golden_rel_l2": orc.rel_l2(o.cpu(), torch.from_numpy(gt["out"]))}
    cfgL = aco.VitConfig()
    t, tsd = tower(cfgL, 1)
    sh, dp = aco.seeded_prompts(cfgL, depth=9, seed=1)
    gen = torch.Generator().manual_seed(3)
    im = torch.randn(2, 3, 336, 336, generator=gen)
    al = (torch.rand(2, 1, 336, 336, generator=gen) > 0.5).float()
    with torch.no_grad():
        o = t(im.cuda(), al.cuda(), sh.cuda(), [d.cuda() for d in dp])
    w = aco.vit_maple_forward(tsd, cfgL, im, al, sh, dp)
    tw["vit_l14_336_rel_l2"] = orc.rel_l2(o.cpu(), w)
    tw["vit_l14_336_min_cosine"] = float(torch.nn.functional.cosine_similarity(o.cpu(), w, dim=-1).min())
    rep["alpha_clip_tower"] = tw
    print("alpha-clip tower:", json.dumps(tw))
    del t
    torch.cuda.empty_cache()
    # how much does the cuFFT high-pass depend on the batch it is computed in?
    cfg = PRESETS["vit_h"]
    x = seeded_images(cfg, 2).cuda()
    a, b = high_pass(x, 0.25)[1:2], high_pass(x[1:2], 0.25)
    rep["high_pass_batch_dependence_max_abs"] = float((a - b).abs().max())
    print("high_pass batch dependence max abs:", rep["high_pass_batch_dependence_max_abs"])
    # ---- BASELINE.json configs[1] north-star lines: fp32 check path at ViT-H, and the mask-IoU probe next to what the
    # REFERENCE ITSELF gives in bf16 (torch.autocast) and in fp32 on this GPU -- the probe's noise floor
    m = build_encoder(cfg).eval()
    m.load_state_dict(seeded_state_dict(cfg), strict=True)
    m = m.cuda()
    gm = np.load(os.path.join(GOLDEN, "encoder_vit_h_multi.npz"))
    stride = int(gm["stride"])
    n_img = 4
    imgs = seeded_images(cfg, n_img, 0).cuda()
    with torch.no_grad():
        chk = torch.cat([m.forward_check(imgs[i:i + 1]) for i in range(n_img)])
        ours = m(imgs)
    ref_sub = torch.from_numpy(gm["out_sub"][:n_img])
    vit_h = {"fp32_check_path_rel_l2_vs_reference": [orc.rel_l2(chk[i, :, ::stride, ::stride].cpu(), ref_sub[i]) for i in range(n_img)],
             "bf16_path_rel_l2_vs_reference": [orc.rel_l2(ours[i, :, ::stride, ::stride].cpu(), ref_sub[i]) for i in range(n_img)]}
    embs = {"ours_bf16": ours.cpu(), "ours_fp32_check_path": chk.cpu()}
    del m
    torch.cuda.empty_cache()
    from oracle import ref_import
    if ref_import.staged_encoder_available():
        ref = ref_import.build_reference_encoder(cfg, staged=True).eval()
        ref.load_state_dict(seeded_state_dict(cfg), strict=True)
        ref = ref.cuda()
        with torch.no_grad():
            r32 = torch.cat([ref(imgs[i:i + 1]) for i in range(n_img)])
            with torch.autocast("cuda", dtype=torch.bfloat16):
                r16 = torch.cat([ref(imgs[i:i + 1]).float() for i in range(n_img)])
        vit_h["reference_gpu_fp32_rel_l2_vs_reference_cpu"] = [orc.rel_l2(r32[i, :, ::stride, ::stride].cpu(), ref_sub[i]) for i in range(n_img)]
        vit_h["reference_autocast_bf16_rel_l2_vs_reference_cpu"] = [orc.rel_l2(r16[i, :, ::stride, ::stride].cpu(), ref_sub[i]) for i in range(n_img)]
        embs["reference_gpu_fp32"] = r32.cpu()
        embs["reference_autocast_bf16"] = r16.cpu()
        del ref
        torch.cuda.empty_cache()
    probe = {}
    for label, e in embs.items():
        per_seed = []
        for sdx in range(gm["logits"].shape[0]):
            dsd_s = mdo.seeded_decoder_state(sdx)
            vals = []
            for i in range(n_img):
                got = mdo.decode_mask_logits(dsd_s, e[i:i + 1])
                want_l = torch.from_numpy(gm["logits"][sdx, i].astype(np.float32))[None, None]
                vals.append(mdo.iou(mdo.binary_mask(got, 1024), mdo.binary_mask(want_l, 1024)))
            per_seed.append({"decoder_seed": sdx, "iou_min": min(vals), "iou_mean": sum(vals) / len(vals)})
        probe[label] = per_seed
    vit_h["mask_iou_probe_vs_reference_cascade"] = probe
    vit_h["note"] = ("IoU of the binary masks of a RANDOM-INIT MaskDecoder_Edge (3 seeds) on each encoder's embedding against the "
                     "reference CPU fp32 cascade, 4 images; the value is set by the decoder draw (how many pixels have a logit "
                     "within the embedding's noise), not by the image")
    rep["vit_h_north_star"] = vit_h
    print("vit_h north star:", json.dumps(vit_h))
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    json.dump(rep, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity.json")
