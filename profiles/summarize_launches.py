"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/summarize_launches.py gpurun_out/launches_r1.csv > profiles/launches_r1.md"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        unit = r["Metric Unit"]
        v = float(r["Metric Value"].replace(",", ""))
        us = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)
        name = r["Kernel Name"]
        name = re.sub(r"\(.*$", "", name)
        rows.append((name, us))
    agg = OrderedDict()
    for n, us in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu launch list summary: {path}\n")
    print(f"{len(rows)} launches, {tot / 1e3:.2f} ms total device time (cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total ms | share | mean us |")
    print("|---|---:|---:|---:|---:|")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        short = n if len(n) < 110 else n[:107] + "..."
        print(f"| `{short}` | {c} | {us / 1e3:.3f} | {100 * us / tot:.1f}% | {us / c:.1f} |")
    mine = sum(a[1] for n, a in agg.items() if "cgl::" in n)
    print(f"\nengine kernels (cgl::*): {100 * mine / tot:.1f}% of device time; "
          f"library/ATen kernels (server-side generator in torch): {100 * (tot - mine) / tot:.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
