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


def _copy(src: str, dst: str, verbose: bool) -> None:
    if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        if verbose:
            print(f"[stage_ref] {src} -> {dst}")


def stage(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref/refsam holds the reference files (freshly copied or already identical).  Also stages
    the files of the whole inference cascade (BASELINE.json configs[3]: wrapper, decoder, Alpha-CLIP, prompt learner,
    tokenizer vocabulary, class names, shipped text features; list in oracle/ref_cascade.py) under oracle/_ref/cascade,
    mirroring the reference tree, for the swapped-vs-not test on the GPU box."""
    if not all(os.path.isfile(os.path.join(SRC_DIR, f)) for f in FILES):
        return all(os.path.isfile(os.path.join(DST_DIR, f)) for f in FILES)
    for f in FILES:
        _copy(os.path.join(SRC_DIR, f), os.path.join(DST_DIR, f), verbose)
    try:
        from oracle.ref_cascade import CASCADE_FILES, STAGED_ROOT
    except ImportError:          # run as a script: oracle/ is sys.path[0]
        from ref_cascade import CASCADE_FILES, STAGED_ROOT
    if all(os.path.isfile(os.path.join(REF_ROOT, f)) for f in CASCADE_FILES):
        for f in CASCADE_FILES:
            _copy(os.path.join(REF_ROOT, f), os.path.join(STAGED_ROOT, f), verbose)
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
