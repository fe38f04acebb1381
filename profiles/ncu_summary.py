#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / bench.py quote.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [ncells]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
ncells = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("---", d.get("Kernel Name", "")[:70], d.get("Grid Size", ""), d.get("Block Size", ""))
    for w in want:
        if w in d:
            print("  %-66s %18s %s" % (w, d[w], units[hdr.index(w)]))
    st = [(float(d[h].replace(",", "")), h) for h in hdr
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and d[h] not in ("", "n/a")]
    for v, h in sorted(st, reverse=True)[:7]:
        print("  stall %-38s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], stderr=subprocess.DEVNULL).decode()
k, hdr2, agg, tot = 0, None, collections.Counter(), 0
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        k += 1; continue
    if r and r[0] == "Address":
        hdr2 = r; continue
    if k != 1 or hdr2 is None or len(r) < 8:
        continue
    d = dict(zip(hdr2, r))
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", d["Source"].strip())
    n = int(d["Instructions Executed"]); agg[m.group(2) if m else "?"] += n; tot += n
print("executed warp instructions (first kernel): %d" % tot + (" = %.0f per warp-cell" % (tot / (ncells / 32)) if ncells else ""))
print("  " + "  ".join("%s %.1f%%" % (op, 100.0 * n / tot) for op, n in agg.most_common(14)))
