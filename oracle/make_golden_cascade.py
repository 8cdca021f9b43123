"""Fixture for BASELINE.json configs[3] (container only, ~2 minutes): the reference's own cascade -- SAM.infer_test through
the reference wrapper + the demo's second CLIP pass (oracle/ref_cascade.py) -- on the synthetic OVCamo-shaped sample, CPU
fp32.  Writes tests/golden/cascade_demo.npz: the final binary mask (bit-packed 1024^2), a strided sample of the mask
logits, the predicted class id and the 61 class scores."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_cascade as rc  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    c = rc.build_cascade("cpu", 0)
    r = c.run(*rc.synthetic_sample(0))
    logits = r["mask_logits"][0, 0]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cascade_demo.npz"),
                        mask_bits=np.packbits((logits > 0).numpy()), logits_sub=logits[::4, ::4].numpy().astype(np.float16),
                        logits_norm=np.float64(logits.double().norm().item()), class_id=r["class_id"].numpy(),
                        score=r["score"].numpy(), class_name=np.array(c.class_names[int(r["class_id"][0])]))
    print("[golden cascade] class", int(r["class_id"][0]), c.class_names[int(r["class_id"][0])], "positive fraction",
          float((logits > 0).float().mean()))


if __name__ == "__main__":
    main()
