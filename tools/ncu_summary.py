"""Summary table of one `ncu --set full` capture of the align kernel -> markdown.
usage: ncu_summary.py <file.ncu-rep> <out.md> "<title line>" ["reading" text file]"""
import csv
import subprocess
import sys

rep, out, title = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size"]
lines = ["# " + title, "", "| metric | value | unit |", "|---|---|---|"]
kernel = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
for i, h in enumerate(hdr):
    keep = h in want
    if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
        try:
            keep = float(vals[i].replace(",", "")) >= 0.2 and "selected_per" not in h.replace("not_selected", "x")
        except ValueError:
            keep = False
    if keep:
        lines.append("| %s | %s | %s |" % (h, vals[i], units[i]))
lines.append("")
lines.append("kernel: `%s`" % kernel)
if len(sys.argv) > 4:
    lines.append("")
    lines.append(open(sys.argv[4]).read().strip())
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:30]))
