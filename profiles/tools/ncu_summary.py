"""Per-launch summary (selected metrics) of an `ncu --set full` report.
usage: ncu_summary.py <report.ncu-rep> <out.csv> "<title line>" """
import csv, subprocess, sys, re
rep, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
cols = [hdr.index(m) for m in WANT]
name = hdr.index("Kernel Name")
lines = [f"# {title}", f"# source: {rep} (not tracked); per launch; cold-cache, serialised",
         "Kernel Name," + ",".join(f"{m} [{units[c]}]" if units[c] else m for m, c in zip(WANT, cols))]
for r in rows[2:]:
    k = re.sub(r"\(.*", "", r[name]).replace("void ", "").replace("mvae::", "").replace("<unnamed>::", "")
    lines.append('"' + k + '",' + ",".join(r[c].replace(",", "") for c in cols))
open(dst, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
