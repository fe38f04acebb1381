#!/usr/bin/env python
"""Top SASS instructions of an `ncu --set full --import-source on` capture by warp-stall samples, with the stall reason that
dominates each and a few instructions of context in front of it (the producer of the register a stalled compare waits for).
usage: python profiles/ncu_sass_hot.py report.ncu-rep [top=12] [context=6]"""
import csv, io, subprocess, sys


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    print(rows[0][1][:160])
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    n = lambda r, k: int(r[ix[k]] or 0)
    tot = sum(n(r, "# Samples") for r in data)
    print("total samples", tot)
    by_reason = {h: sum(n(r, h) for r in data) for h in stall_cols}
    print("by reason:", "  ".join("%s %.1f%%" % (h[6:], 100.0 * v / tot) for h, v in sorted(by_reason.items(), key=lambda kv: -kv[1])[:8]))
    top = sorted(range(len(data)), key=lambda i: -n(data[i], "# Samples"))[:top_n]
    for i in sorted(top):
        r = data[i]
        why = max(stall_cols, key=lambda h: n(r, h))
        print("--- #%d  %.1f%% of samples, mostly %s" % (i, 100.0 * n(r, "# Samples") / tot, why[6:]))
        for j in range(max(0, i - ctx), i + 1):
            print("   %5d %6d  %s" % (j, n(data[j], "# Samples"), data[j][ix["Source"]][:120]))


if __name__ == "__main__":
    main()
