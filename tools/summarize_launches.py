"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name, shares."""
import csv
import re
import sys
from collections import defaultdict


def main(path, last_n=None):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((int(r["ID"]), name, ns))
    if last_n:
        rows = rows[-last_n:]
    agg = defaultdict(lambda: [0, 0.0])
    for _, name, ns in rows:
        agg[name][0] += 1
        agg[name][1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"launches={len(rows)} total={total / 1e3:.1f} us")
    for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns / 1e3:10.1f} us {100 * ns / total:5.1f}%  x{cnt:<4d} {name[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
