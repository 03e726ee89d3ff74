"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time and count per kernel."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, title=""):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rd:
        name = re.sub(r"\(.*", "", r[ki])
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
        a = agg.setdefault(name, [0.0, 0])
        a[0] += us
        a[1] += 1
    tot = sum(a[0] for a in agg.values())
    n = sum(a[1] for a in agg.values())
    if title:
        print(title)
    for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{t:9.1f} us  x {c:<3d} {100 * t / tot:5.1f}%  {k}")
    print(f"total {tot:.1f} us over {n} launches")


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
