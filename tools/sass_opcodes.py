#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libcvlm_b200.so (no GPU needed): how many tcgen05 MMA (UTCHMMA / UTCQMMA ...),
tensor-memory loads / stores (LDTM / STTM), TMA loads / stores (UTMALDG / UTMASTG / UBLKCP), tcgen05 barriers
(UTCBAR) and legacy HMMA instructions each kernel contains -- the evidence that the hot kernels are tcgen05 + TMA code.

    python tools/sass_opcodes.py [out.json]          # default: profiles/sass_opcodes_r2.json
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "camouflaged_vlm_b200", "libcvlm_b200.so")
WATCH = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UTCBAR",
         "UTCCP", "HMMA", "SYNCS", "ELECT", "MUFU", "FFMA2", "FADD2", "FMUL2", "LDG", "STG", "LDS", "STS", "BAR")


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_opcodes_r2.json")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P[T\d]+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            kernels[cur]["_total"] += 1
            op = m.group(1)
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[cur][w] += 1
                    break
            else:
                if op.startswith("UTC") or op.startswith("UTMA"):
                    kernels[cur][op] += 1
    names = subprocess.run(["cu++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    report = {"library": os.path.relpath(LIB, ROOT), "arch": "sm_100a", "kernels": {}}
    total = collections.Counter()
    for mangled, pretty in zip(kernels, names):
        c = kernels[mangled]
        short = pretty.replace("cvlm::", "").replace("void ", "", 1)
        short = short[:short.index(">(") + 1] if ">(" in short else re.sub(r"\(.*", "", short)
        key = short if short not in report["kernels"] else mangled
        report["kernels"][key] = {k: v for k, v in sorted(c.items()) if v}
        total.update(c)
    report["total"] = {k: v for k, v in sorted(total.items()) if v}
    with open(out_path, "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report["total"]))
    for k, c in report["kernels"].items():
        if c.get("UTCHMMA") or c.get("UTMALDG") or c.get("UTMASTG"):
            print(f"{k[:100]:100s} " + " ".join(f"{w}={c[w]}" for w in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA") if c.get(w)))


if __name__ == "__main__":
    main()
