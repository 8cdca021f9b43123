"""Per-region warp-stall sample summary of an `ncu --set full --import-source on` report (read here, no GPU).

    python tools/ncu_regions.py gpurun_out/prof.ncu-rep [top_n]

Groups consecutive SASS instructions by execution count (a cheap proxy for "same loop / same warp role"),
prints the samples per group and the hottest instructions with their dominant stall reasons."""
import csv
import io
import subprocess
import sys


def main(rep, top_n=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, data = rows[1], rows[2:]
    isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[isamp]) for r in data)
    print("total samples", tot, "instructions", len(data))

    def s(a, b):
        return sum(int(data[i][isamp]) for i in range(a, b))

    prev, start = None, 0
    for i, r in enumerate(data + [None]):
        ex = int(r[iex]) if r is not None else None
        if ex != prev:
            if prev is not None and s(start, i) > tot * 0.003:
                print(f"  region {start:5d}-{i:5d}  executed {prev:9d}  samples {s(start, i):7d}  ({100.0 * s(start, i) / tot:4.1f}%)")
            prev, start = ex, i
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top_n]
    for i in sorted(idx):
        r = data[i]
        st = {h[c][6:]: int(r[c]) for c in stall_cols if int(r[c]) > 0}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f"{i:5d} {r[isrc].strip()[:64]:64s} {int(r[isamp]):6d} {int(r[iex]):9d} {top}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
