"""per-line and per-function view of an ncu report's source page (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py <report.ncu-rep> [kernel-id] [top-n]     prints the top source lines by warp-instructions
with lanes per instruction and stall samples, per file"""
import collections, csv, os, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
args = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 2 and sys.argv[2] != "-":
    args += ["--launch-skip", sys.argv[2], "--launch-count", "1"]
rows = list(csv.reader(subprocess.run(args, capture_output=True, text=True).stdout.splitlines()))
cur = None
lines = []
tot = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0].isdigit() and r[2] == "-":
        w, t, s = int(r[7]), int(r[8]), int(r[6])
        tot += w
        lines.append((w, t, s, cur, int(r[0]), r[1].strip()))
lines.sort(key=lambda x: -x[0])
print("total warp-instructions: %d" % tot)
print("%10s %6s %5s %7s  %s" % ("warp_inst", "share", "lanes", "samples", "file:line  source"))
for w, t, s, f, ln, src in lines[:top]:
    print("%10d %5.1f%% %5.1f %7d  %s:%d  %s" % (w, 100.0 * w / max(tot, 1), t / max(w, 1), s, f, ln, src[:110]))
