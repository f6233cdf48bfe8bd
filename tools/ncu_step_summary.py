"""One step's kernels from an `ncu --set full` report (tools/ncu_two.sh) -> markdown table + profiles/traffic.json.
usage: ncu_step_summary.py <file.ncu-rep> <out.md> <traffic.json> "<title>" """
import csv, json, subprocess, sys
rep, out_md, out_json, title = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "registers"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue")]
def num(v):
    try: return float(v.replace(",", ""))
    except ValueError: return None
def scaled(v, u):
    x = num(v)
    if x is None: return v
    f = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9}.get(u)
    return x * f if f else x
kern = []
for r in rows[2:]:
    if len(r) != len(hdr): continue
    name = r[ix["Kernel Name"]]
    d = {"name": name}
    for m, lab in want:
        if m in ix: d[lab] = scaled(r[ix[m]], units[ix[m]])
    kern.append(d)
lines = ["# " + title, "", "| metric | " + " | ".join("`%s` #%d" % (k["name"].split("(")[0][-28:], i) for i, k in enumerate(kern)) + " |", "|---|" + "---|" * len(kern)]
for m, lab in want:
    vals = []
    for k in kern:
        v = k.get(lab)
        if isinstance(v, float):
            vals.append("%.3f ms" % (v * 1e3) if lab == "time" else "%.2f GB" % (v / 1e9) if lab.startswith("DRAM r") or lab.startswith("DRAM w") else "%.3g" % v)
        else: vals.append(str(v))
    lines.append("| %s | %s |" % (lab, " | ".join(vals)))
tot_t = sum(k.get("time") or 0 for k in kern); tot_dram = sum((k.get("DRAM read") or 0) + (k.get("DRAM write") or 0) for k in kern)
tot_inst = sum(k.get("warp instructions") or 0 for k in kern)
lines += ["", "step total (under ncu, serialised, cold caches): %.3f ms, %.2f GB of DRAM traffic, %.3g warp instructions" % (tot_t * 1e3, tot_dram / 1e9, tot_inst)]
open(out_md, "w").write("\n".join(lines) + "\n")
json.dump({"dram_bytes_per_launch": tot_dram, "source": rep.split("/")[-1],
           "executed": {"warp_instructions_per_launch": tot_inst, "thread_instruction_slots_per_launch": tot_inst * 32,
                        "per_kernel": [{"kernel": k["name"].split("(")[0], "ms": (k.get("time") or 0) * 1e3, "warp_instructions": k.get("warp instructions"),
                                        "issue_slots_busy_pct": k.get("issue slots busy %"), "alu_pipe_pct": k.get("ALU pipe %"),
                                        "dram_bytes": (k.get("DRAM read") or 0) + (k.get("DRAM write") or 0)} for k in kern]}}, open(out_json, "w"), indent=1)
print("\n".join(lines))
