"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.  usage: ncu_launch_table.py file.csv [passes]"""
import csv
import sys
from collections import defaultdict


def main(path, passes=1):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    h = rows[0]
    ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    ui = h.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        a = agg[r[ki][:70]]
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | us / pass | share |\n|---|---:|---:|---:|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {us / passes:.1f} | {us / tot * 100:.1f}% |")
    print(f"| **total** | | {tot / passes:.1f} | 100% |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
