"""Golden vectors for the image input pipeline (datasets/wrappers.py:22-27): PIL + torchvision, the reference's own
third-party implementation, run in this container.  Writes tests/golden/preprocess.npz.

    python oracle/make_golden_preprocess.py
"""
import os
import sys

import numpy as np
from PIL import Image
from torchvision import transforms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [   # (h, w, out_size): up-scale, down-scale (wide kernels), mixed, identity, extreme aspect
    (37, 53, 64), (150, 100, 64), (64, 64, 64), (90, 333, 96), (301, 17, 48), (64, 200, 64),
]


def main():
    rng = np.random.RandomState(20250101)
    out = {"cases": np.asarray(CASES, np.int32)}
    for i, (h, w, s) in enumerate(CASES):
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        # smooth half of the cases so that the resample is not only averaging noise
        if i % 2:
            yy, xx = np.mgrid[0:h, 0:w]
            img = ((np.sin(yy[..., None] / 7.0 + np.arange(3)) * np.cos(xx[..., None] / 11.0) * 0.5 + 0.5) * 255).astype(np.uint8)
        pil = Image.fromarray(img)
        t = transforms.Compose([transforms.Resize((s, s)), transforms.ToTensor(),
                                transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        out[f"in{i}"] = img
        out[f"u8_{i}"] = np.asarray(pil.resize((s, s), Image.BILINEAR))
        out[f"out{i}"] = t(pil).numpy()
    path = os.path.join(ROOT, "tests", "golden", "preprocess.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; PIL", Image.__version__ if hasattr(Image, "__version__") else "")


if __name__ == "__main__":
    sys.exit(main())
