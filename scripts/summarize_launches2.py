"""Launch list with several metrics per launch (`ncu --metrics a,b,c --csv`, long format) -> markdown: per-kernel totals
and EVERY launch of one steady-state step with its device time, tensor-pipe activity, DRAM and L2 bytes.
    python scripts/summarize_launches2.py gpurun_out/launches_r02_f16x3.csv "title" > profiles/launches_r02_f16x3.md"""
import csv
import re
import sys
from collections import OrderedDict, defaultdict

UNIT = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0, "": 1.0}


def main(path, title):
    lines = [l for l in open(path, newline="") if not l.startswith("==")]
    launches = OrderedDict()
    for r in csv.DictReader(lines):
        i = int(r["ID"])
        d = launches.setdefault(i, {"name": r["Kernel Name"], "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        d[r["Metric Name"]] = v * UNIT.get(r.get("Metric Unit", ""), 1.0)
    T, TP = "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    DR, DW, L2, WA = "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"
    short = lambda n: re.sub(r"\(.*", "", n).replace("yb::<unnamed>::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("yb::", "").replace("void ", "").strip()
    total = sum(d.get(T, 0.0) for d in launches.values())
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
    for d in launches.values():
        a = agg[re.sub(r"<.*", "", short(d["name"]))]
        a[0] += 1
        a[1] += d.get(T, 0.0)
        a[2] += d.get(DR, 0.0) + d.get(DW, 0.0)
        a[3] += d.get(L2, 0.0)
        a[4] += d.get(TP, 0.0) * d.get(T, 0.0)
    print("# " + title + "\n")
    print("%d launches, %.3f ms summed device time (ncu: serialised, cold cache -- compare SHARES, not absolutes)\n" % (len(launches), total / 1e6))
    print("| kernel | launches | total ms | share | DRAM MB | L2 MB | time-weighted tensor pipe % |\n|---|---:|---:|---:|---:|---:|---:|")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% | %.0f | %.0f | %.1f |" % (n, a[0], a[1] / 1e6, 100 * a[1] / total, a[2] / 1e6, a[3] / 1e6, a[4] / a[1] if a[1] else 0))
    print("\n## every launch\n\n| id | kernel | grid | block | us | tensor pipe % | warps active % | DRAM MB | L2 MB |\n|---:|---|---|---|---:|---:|---:|---:|---:|")
    for i, d in launches.items():
        print("| %d | `%s` | %s | %s | %.1f | %.1f | %.1f | %.1f | %.1f |" % (i, short(d["name"])[:70], d["grid"], d["block"], d.get(T, 0) / 1e3, d.get(TP, 0), d.get(WA, 0),
                                                                        (d.get(DR, 0) + d.get(DW, 0)) / 1e6, d.get(L2, 0) / 1e6))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
