#!/usr/bin/env python
"""Reduce an `ncu --set full` report to the handful of per-launch numbers DESIGN.md / bench.py cite.

    python scripts/ncu_summary.py gpurun_out/X.ncu-rep profiles/X_ncu_full_summary.csv
    python scripts/ncu_summary.py --traffic profiles/traffic.json c2:4096:profiles/A_summary.csv c4:8192:profiles/B_summary.csv ...
        (writes the dram__bytes_read + dram__bytes_write per launch that bench.py prints as roofline.traffic, with its source)
"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_bytes.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max"]


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def traffic(out, specs):
    import json
    import os
    res = {}
    for spec in specs:
        wl, batch, path = spec.split(":", 2)
        rows = list(csv.reader(open(path)))
        hdr, units, first = rows[0], rows[1], rows[2]
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(key)
            tot += float(first[i]) * UNIT[units[i]]
        res[wl] = {"bytes": tot, "batch": int(batch), "kernel": first[hdr.index("Kernel Name")].split("(")[0],
                   "source": path if not os.path.isabs(path) else os.path.relpath(path)}
    json.dump(res, open(out, "w"), indent=1, sort_keys=True)
    print("wrote", out, {k: v["bytes"] for k, v in res.items()})


def main():
    if sys.argv[1] == "--traffic":
        return traffic(sys.argv[2], sys.argv[3:])
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEYS if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    print("wrote", out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main()
