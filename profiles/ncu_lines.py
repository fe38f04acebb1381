#!/usr/bin/env python
"""Per-source-line executed warp instructions and stall samples of the first kernel in an .ncu-rep.
usage: python profiles/ncu_lines.py rep [top_n]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stderr=subprocess.DEVNULL).decode()
hdr = None; inst = collections.Counter(); samp = collections.Counter(); text = {}; nk = 0
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Function Name":
        nk += 1
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or nk != 1 or len(r) < 8: continue
    # cuda,sass view: rows with a line number are CUDA lines followed by their SASS rows (empty line no)
    d = dict(zip(hdr, r))
    ln = r[0]
    if ln:
        cur = int(ln); text[cur] = r[1]
        try:
            inst[cur] += int(r[hdr.index("Instructions Executed")] or 0); samp[cur] += int(r[hdr.index("# Samples")] or 0)
        except ValueError:
            pass
ti = sum(inst.values()) or 1; ts = sum(samp.values()) or 1
print("total warp instructions %d, samples %d" % (ti, ts))
for ln, n in inst.most_common(top):
    print("%5d %6.2f%% inst %6.2f%% samp | %s" % (ln, 100.0 * n / ti, 100.0 * samp[ln] / ts, text[ln].strip()[:110]))
