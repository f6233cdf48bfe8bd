"""Per-source-line stall samples of one kernel from an ncu report (sass,cuda source page).
usage: src_hot.py file.ncu-rep [top N]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "-"):
        ix = {h: i for i, h in enumerate(hdr)}
        def num(x):
            try: return float(x.replace(",", ""))
            except ValueError: return 0.0
        lines.append((num(r[ix["# Samples"]]), num(r[ix["Instructions Executed"]]), cur, r[0], r[1].strip()[:110],
                      num(r[ix["stall_long_sb"]]), num(r[ix["stall_short_sb"]]), num(r[ix["stall_wait"]]), num(r[ix["stall_math"]]), num(r[ix["stall_lg"]])))
tot = sum(l[0] for l in lines); toti = sum(l[1] for l in lines)
print("samples %d, warp instructions %d" % (tot, toti))
for l in sorted(lines, key=lambda l: -l[0])[:top]:
    print("%5.2f%% inst %5.2f%% lsb %5.0f ssb %5.0f wait %5.0f math %5.0f lg %4.0f  %s:%s  %s" % (100 * l[0] / tot, 100 * l[1] / toti, l[5], l[6], l[7], l[8], l[9], l[2], l[3], l[4]))
