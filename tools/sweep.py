"""BASELINE.json configs[4]: encoder shape sweep -- ViT-B/L/H at 1024^2, batch 1..128, window 14 vs
global-only attention.  Prints one JSON line per point (img/s, ms, achieved algorithmic TFLOP/s).  GPU only.
Usage: python tools/sweep.py [out.jsonl] [max_batch]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from camouflaged_vlm_b200 import PRESETS, build_encoder, global_only


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep.jsonl"
    max_batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(0)
    rows = []
    for name in ("vit_b", "vit_l", "vit_h"):
        for variant in ("window14", "global_only"):
            cfg = PRESETS[name] if variant == "window14" else global_only(PRESETS[name])
            model = build_encoder(cfg).eval().to(dev)
            with torch.no_grad():
                for n, p in model.named_parameters():
                    p.copy_(torch.randn(p.shape, generator=gen, device=dev) * 0.02)
                    if n.endswith(("norm1.weight", "norm2.weight", "neck.1.weight", "neck.3.weight")):
                        p.add_(1.0)
            for batch in (1, 2, 4, 8, 16, 32, 64, 128):
                if batch > max_batch:
                    break
                x = torch.randn(batch, 3, cfg.img_size, cfg.img_size, generator=gen, device=dev)
                with torch.no_grad():
                    for _ in range(2):
                        model(x)
                    torch.cuda.synchronize()
                    iters = 3 if batch >= 32 else 5
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(iters):
                        model(x)
                    e1.record()
                    torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                row = {"model": name, "attention": variant, "batch": batch, "ms": ms, "img_per_s": batch / ms * 1e3,
                       "tflops": batch * cfg.flops_per_image() / ms / 1e9}
                rows.append(row)
                print(json.dumps(row), flush=True)
                del x
            del model
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    with open(out_path, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
