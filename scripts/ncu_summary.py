"""ncu report -> markdown table of the metrics the roofline discussion uses.
   python scripts/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/x.md"""
import csv, io, re, subprocess, sys
METRICS = ["gpu__time_duration.sum", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum",
           "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__shared_mem_per_block_dynamic", "launch__registers_per_thread", "launch__cluster_size"]
rep, title = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = ["ID", "Kernel Name", "Grid Size", "Block Size"] + [m for m in METRICS if m in hdr]
idx = [hdr.index(c) for c in cols]
print("# " + title + "\n")
print("| " + " | ".join(cols) + " |")
print("|" + "---|" * len(cols))
print("| " + " | ".join(units[i] for i in idx) + " |")
for r in data:
    vals = [r[i] for i in idx]
    m = re.search(r"(\w+_kernel<[^>]*>|\w+_kernel)", vals[1])
    vals[1] = "`%s`" % (m.group(1) if m else vals[1][:60])
    print("| " + " | ".join(vals) + " |")
