#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel:  python profiles/summarize_launches.py launches.csv"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:70]


def main(path):
    tot, cnt = defaultdict(float), defaultdict(int)
    with open(path) as f:
        rows = [r for r in f if r.startswith('"')]
    for row in csv.DictReader(rows):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short(row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"] == "us":
            v *= 1e3
        elif row["Metric Unit"] == "ms":
            v *= 1e6
        tot[k] += v / 1e3
        cnt[k] += 1
    total = sum(tot.values())
    print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for k in sorted(tot, key=lambda k: -tot[k])[:24]:
        print(f"| `{k}` | {cnt[k]} | {tot[k]:.1f} | {tot[k] / total:.3f} | {tot[k] / cnt[k]:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
