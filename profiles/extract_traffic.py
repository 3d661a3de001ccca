#!/usr/bin/env python
"""Per-kernel DRAM traffic and key counters from one `ncu --set full` capture:
    ncu -i capture.ncu-rep --page raw --csv > raw.csv ; python profiles/extract_traffic.py raw.csv [out.json]
Writes {kernel: {dram_read_bytes, dram_write_bytes, time_us, sm_throughput_pct, issue_active_pct, warps_active_pct, registers,
inst_executed}} (first instance of every kernel; per launch).  bench.py reads profiles/traffic.json for `roofline.traffic`."""
import csv
import json
import re
import sys

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
KEYS = {
    "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes", "gpu__time_duration.sum": "time_us",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers", "smsp__inst_executed.sum": "inst_executed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
}


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^mb::", "", name)
    return re.sub(r"<.*$", "", name).replace("_kernel", "")


def main(path, out=None):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    res = {}
    for d in data:
        k = short(d[col["Kernel Name"]])
        if k in res:
            continue
        e = {}
        for m, nm in KEYS.items():
            if m in col and d[col[m]] not in ("", "n/a"):
                e[nm] = float(d[col[m]].replace(",", "")) * UNITS.get(units[col[m]], 1.0)
        e["grid"] = d[col["Grid Size"]] if "Grid Size" in col else None
        res[k] = e
    txt = json.dumps(res, indent=1, sort_keys=True)
    if out:
        open(out, "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
