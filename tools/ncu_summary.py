"""Summarise `ncu --set full` reports (read here, no GPU needed) into small JSON files under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_gemm_r1.ncu-rep profiles/ncu_gemm_r1.json [batch]
"""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__cycles_elapsed.avg": "sm_cycles",
    "smsp__inst_executed.sum": "warp_instructions",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}


def main(rep, out, batch=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for k, name in KEEP.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[name] = v * UNIT_SCALE.get(units[i], 1.0)
        if "dram_read" in d and "dram_write" in d:
            d["dram_bytes"] = d["dram_read"] + d["dram_write"]
        launches.append(d)
    summ = {"report": rep, "batch": batch, "launches": launches}
    if launches and all("dram_bytes" in l for l in launches):
        summ["mean_dram_bytes_per_launch"] = sum(l["dram_bytes"] for l in launches) / len(launches)
    json.dump(summ, open(out, "w"), indent=1)
    for l in launches:
        print({k: (round(v, 6) if isinstance(v, float) else v) for k, v in l.items()})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)
