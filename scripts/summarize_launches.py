"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table
(per-kernel share of the step + the slowest individual launches).  Usage:
    python scripts/summarize_launches.py gpurun_out/launches_r01.csv > profiles/launches_r01.md"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        name = name.replace("yb::<unnamed>::", "").replace("void ", "")
        rows.append((name, ns, r.get("Grid Size", ""), r.get("Block Size", ""), int(r["ID"])))
    total = sum(r[1] for r in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns, *_ in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    print("# ncu launch list summary: `%s`\n" % path)
    print("%d launches, %.3f ms summed device time (serialised, cold-cache: compare SHARES, not absolutes)\n" % (len(rows), total / 1e6))
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% |" % (n, c, t / 1e6, 100 * t / total))
    print("\n## 25 slowest launches\n\n| id | kernel | grid | block | us |\n|---:|---|---|---|---:|")
    for n, ns, g, b, i in sorted(rows, key=lambda r: -r[1])[:25]:
        print("| %d | `%s` | %s | %s | %.1f |" % (i, n, g, b, ns / 1e3))


if __name__ == "__main__":
    main(sys.argv[1])
