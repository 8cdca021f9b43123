"""Stage the reference's own implementation of the hot path under oracle/_ref/ (git-ignored, travels to the GPU box).

    python oracle/stage_ref.py            # run by __graft_entry__.build() whenever /root/reference is present

The reference is Python: its encoder lives in two torch-only files (models/mmseg/models/sam/image_encoder.py and
common.py; image_encoder.py:7-22 imports nothing else).  "Building" it is placing those files, UNMODIFIED, where
the GPU box can import them: oracle/_ref/refsam/.  With them `bench.py --impl reference` and the `cpu_baseline` /
`torch_gpu_baseline` legs time the reference module itself (``cpu_baseline.kind = "reference"``) instead of the oracle
port.  Nothing here is product code and nothing under camouflaged_vlm_b200/ imports it; oracle/_ref/ is never
committed (.gitignore), the sources stay the reference's.
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.environ.get("CVLM_REFERENCE_ROOT", "/root/reference")
SRC_DIR = os.path.join(REF_ROOT, "models", "mmseg", "models", "sam")
DST_DIR = os.path.join(ROOT, "oracle", "_ref", "refsam")
FILES = ("image_encoder.py", "common.py")


def stage(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref/refsam holds the reference files (freshly copied or already identical)."""
    if not all(os.path.isfile(os.path.join(SRC_DIR, f)) for f in FILES):
        return all(os.path.isfile(os.path.join(DST_DIR, f)) for f in FILES)
    os.makedirs(DST_DIR, exist_ok=True)
    for f in FILES:
        src, dst = os.path.join(SRC_DIR, f), os.path.join(DST_DIR, f)
        if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            if verbose:
                print(f"[stage_ref] {src} -> {dst}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
