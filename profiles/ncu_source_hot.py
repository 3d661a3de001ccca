#!/usr/bin/env python
"""Print the SASS listing of one kernel from `ncu --page source --csv` with samples, executed counts and dominant stall:
python profiles/ncu_source_hot.py source.csv [min_samples [instance]]"""
import csv
import sys


def main(path, min_samples=0, instance=0):
    rows = list(csv.reader(open(path)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    rows = rows[heads[instance] - 1:]          # the instance-th kernel of the export
    h = 1
    hdr = rows[h]
    end = next((i for i in range(h + 1, len(rows)) if rows[i] and rows[i][0] in ("Address", "Kernel Name")), len(rows))
    rows = rows[:end]
    c_src, c_s, c_ie = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    c_thr = hdr.index("Avg. Predicated-On Threads Executed")
    stalls = [(i, n) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
    tot_s = sum(int(r[c_s] or 0) for r in rows[h + 1:] if len(r) > c_s)
    tot_i = sum(int(r[c_ie] or 0) for r in rows[h + 1:] if len(r) > c_ie)
    print(f"total samples {tot_s}, total warp instructions {tot_i}")
    for k, r in enumerate(rows[h + 1:]):
        if len(r) <= c_s:
            continue
        s = int(r[c_s] or 0)
        if s < min_samples:
            continue
        top = sorted(((int(r[i] or 0), n) for i, n in stalls), reverse=True)[:2]
        tops = " ".join(f"{n[6:]}:{v}" for v, n in top if v)
        print(f"{k:5d} {s:6d} {100 * s / max(tot_s, 1):5.1f}% {int(r[c_ie] or 0):9d} {r[c_thr]:>5} {r[c_src].strip()[:90]:90s} {tops}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
