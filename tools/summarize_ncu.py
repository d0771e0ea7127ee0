#!/usr/bin/env python3
"""Developer tool: turns ncu exports into the small text summaries kept under profiles/.

    summarize_ncu.py launches <launches.csv>               per-kernel totals / shares of a launch list
                                                            (ncu --metrics gpu__time_duration.sum --csv)
    summarize_ncu.py raw <report.ncu-rep> [kernel-regex]   key metrics + top stall reasons per launch
                                                            (ncu --set full), via `ncu -i ... --page raw --csv`
    summarize_ncu.py source <report.ncu-rep> <kernel-regex> [top]   hottest SASS instructions
"""
import collections
import csv
import io
import re
import subprocess
import sys

RAW_METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).strip()
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit == "ns" else (v / 1e3 if unit.startswith("us") else (v if unit == "ms" else v * 1e3))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print("%-64s %6s %11s %10s %7s" % ("kernel", "n", "total ms", "avg ms", "share"))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-64s %6d %11.3f %10.4f %7.3f" % (k[:64], c, t, t / c, t / tot))
    print("%-64s %6d %11.3f" % ("TOTAL (cold-cache, serialised: compare shares, not absolutes)", sum(a[0] for a in agg.values()), tot))


def _export(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw(rep, pattern=None):
    rows = _export(rep, "raw")
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if pattern and not re.search(pattern, name):
            continue
        print("---- " + re.sub(r"\(.*", "", name)[:90])
        print("   " + "  ".join("%s=%s%s" % (short, r[idx[m]], units[idx[m]].replace("byte", "B").replace("second", "s"))
                                for m, short in RAW_METRICS if m in idx))
        top = sorted(((float(r[idx[h]].replace(",", "")), h) for h in stalls), reverse=True)[:6]
        print("   stalls/issue: " + ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace(
            "_per_issue_active.ratio", ""), v) for v, h in top))


def source(rep, pattern, top=40):
    rows = _export(rep, "source", ["--kernel-name", "regex:" + pattern, "--launch-count", "1"])
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > idx["# Samples"] and r[idx["# Samples"]].isdigit()]
    half = len(data) // 2 if len(data) > 2 and data[0][idx["Source"]] == data[len(data) // 2][idx["Source"]] else len(data)
    data = data[:half]
    tot = sum(int(r[idx["# Samples"]]) for r in data)
    print("%d SASS instructions, %d stall samples" % (len(data), tot))
    hot = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[:top]
    for i in sorted(hot):
        r = data[i]
        print("%5d %7s %5.1f%%  %s" % (i, r[idx["# Samples"]], 100.0 * int(r[idx["# Samples"]]) / max(tot, 1), r[idx["Source"]].strip()[:100]))


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2])
    elif cmd == "raw":
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    elif cmd == "source":
        source(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
