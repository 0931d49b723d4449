"""Turns the raw ncu outputs brought back in gpurun_out/ into the small text/JSON summaries committed under
profiles/.  Usage: python profiles/summarize.py <round-tag>   (reads gpurun_out/launches_<tag>.csv,
gpurun_out/prof_kstep_<tag>.ncu-rep, gpurun_out/prof_krollout_<tag>.ncu-rep)"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")

RAW_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def launches(tag):
    p = os.path.join(G, "launches_%s.csv" % tag)
    rows = list(csv.reader(open(p)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, mi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 2:]:
        if len(r) <= mi:
            continue
        k = re.sub(r"\(.*", "", r[ki])
        agg.setdefault(k, []).append(float(r[mi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    lines = ["# ncu launch list (gpu__time_duration.sum, --clock-control none) of `python bench.py --steps 20 --warmup 3 --no-cpu`",
             "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes", "",
             "%-34s %6s %14s %12s %7s" % ("kernel", "n", "sum_us", "avg_us", "share")]
    for k, v in agg.items():
        lines.append("%-34s %6d %14.1f %12.1f %6.1f%%" % (k[:34], len(v), sum(v) / 1e3, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    open(os.path.join(OUT, "launches_%s.txt" % tag), "w").write("\n".join(lines) + "\n")
    return agg


def raw(rep):
    rows = list(csv.reader(ncu(["-i", rep, "--page", "raw", "--csv"]).splitlines()))
    H, U = rows[0], rows[1]
    out = {}
    for m in RAW_METRICS:
        if m in H:
            i = H.index(m)
            out[m] = {"unit": U[i], "values": [r[i] for r in rows[2:]]}
    for i, h in enumerate(H):
        if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h:
            out[h] = {"unit": U[i], "values": [r[i] for r in rows[2:]]}
    return out


def by_function(rep):
    core = open(os.path.join(ROOT, "pomcpp_b200", "csrc", "pom_core.cuh")).read().splitlines()
    fn = []
    for i, l in enumerate(core, 1):
        m = re.match(r"^POM_HD(?:_COLD)?\s+[\w:&\*<> ]+?\s+(\w+)\(", l)
        if m:
            fn.append((i, m.group(1)))

    def func_of(line):
        name = "?"
        for i, n in fn:
            if i <= line:
                name = n
            else:
                break
        return name
    rows = list(csv.reader(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]).splitlines()))
    cur, acc = None, {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0].isdigit() and r[2] == "-":
            k = func_of(int(r[0])) if cur == "pom_core.cuh" else cur
            a = acc.setdefault(k, [0, 0, 0])
            a[0] += int(r[7]); a[1] += int(r[8]); a[2] += int(r[6])
    tot = sum(a[0] for a in acc.values()) or 1
    lines = ["%-26s %12s %7s %6s %8s" % ("function", "warp_inst", "share", "lanes", "samples")]
    for k, (w, t, s) in sorted(acc.items(), key=lambda x: -x[1][0])[:30]:
        lines.append("%-26s %12d %6.1f%% %6.1f %8d" % (k, w, 100 * w / tot, t / max(w, 1), s))
    return lines


def by_line(rep, top=45):
    """top source lines by warp-instructions (tools/ncu_lines.py)"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "0", str(top)],
                         capture_output=True, text=True).stdout
    return out.splitlines()


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    if os.path.exists(os.path.join(G, "launches_%s.csv" % tag)):
        launches(tag)
    for kern in ("kstep", "krollout", "kexpand", "kobs"):
        rep = os.path.join(G, "prof_%s_%s.ncu-rep" % (kern, tag))
        if not os.path.exists(rep):
            continue
        r = raw(rep)
        summ = {"source": "ncu --set full --clock-control none --import-source on, " + os.path.basename(rep), "metrics": r}
        if kern in ("kstep", "kexpand", "kobs"):
            rd = [float(v.replace(",", "")) for v in r["dram__bytes_read.sum"]["values"]]
            wr = [float(v.replace(",", "")) for v in r["dram__bytes_write.sum"]["values"]]
            unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
            summ["dram_bytes_per_launch"] = (sum(rd) * unit[r["dram__bytes_read.sum"]["unit"]] +
                                             sum(wr) * unit[r["dram__bytes_write.sum"]["unit"]]) / len(rd)
        if kern == "kstep":
            # r1 captured whole-batch launches of k_step (1 Mi envs); r2 captures k_step_ws in the bench's POM_STEP_OVERLAP
            # phase, where one launch steps half of the batch
            envs = (1 << 20) if tag == "r1" else (1 << 19)
            summ["envs_per_launch"] = envs
            summ["algorithmic_bytes_per_launch"] = 582 * envs
        if kern == "kexpand":
            summ["children_per_launch"] = 4096 * 1296
            summ["algorithmic_bytes_per_launch"] = 289 * 4096 * 1296
        if kern == "kobs":
            summ["envs_per_launch"] = 1 << 20
            summ["algorithmic_bytes_per_launch"] = (289 + 496) * (1 << 20)
        json.dump(summ, open(os.path.join(OUT, "k_%s_ncu_summary%s.json" % (kern[1:], "" if tag == "r1" and kern == "kstep" else "_" + tag)), "w"), indent=1)
        lines = by_function(rep) + ["", "# top source lines"] + by_line(rep)
        open(os.path.join(OUT, "k_%s_by_function_%s.txt" % (kern[1:], tag)), "w").write("\n".join(lines) + "\n")
    print("profiles written for", tag)


if __name__ == "__main__":
    main()
