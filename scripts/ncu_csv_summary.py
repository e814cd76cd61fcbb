"""`ncu -i X.ncu-rep --page raw --csv` output (exported on the GPU box by scripts/gpu_r2_profiles.sh) -> markdown table of
the metrics the roofline discussion uses.   python scripts/ncu_csv_summary.py gpurun_out/prof_X_raw.csv "title" > profiles/x.md"""
import csv
import re
import sys

METRICS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
           "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed.sum", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic",
           "launch__registers_per_thread", "launch__cluster_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(path, title):
    rows = list(csv.reader(open(path, newline="")))
    rows = [r for r in rows if r and not r[0].startswith("==")]
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = ["ID", "Kernel Name", "Grid Size", "Block Size"] + [m for m in METRICS if m in hdr]
    idx = [hdr.index(c) for c in cols]
    print("# " + title + "\n")
    print("ncu `--set full --clock-control none`, one row per captured launch (cold cache, serialised: read SHARES and "
          "percentages, not absolute times).\n")
    print("| " + " | ".join(cols) + " |")
    print("|" + "---|" * len(cols))
    print("| " + " | ".join(units[i] for i in idx) + " |")
    for r in data:
        vals = [r[i] for i in idx]
        m = re.search(r"(\w+_kernel)(<[^>]*>)?", vals[1])
        vals[1] = "`%s%s`" % (m.group(1), m.group(2) or "") if m else "`%s`" % vals[1][:60]
        print("| " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
