#!/usr/bin/env python
"""print registers / spills / smem of every kernel from build/*.ptxas.log (nvcc -Xptxas -v output)"""
import glob, os, re, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info    : Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info    : Used (\d+) registers(.*)")
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for f in sorted(glob.glob(os.path.join(here, "build", "*.ptxas.log"))):
    for m in pat.finditer(open(f).read()):
        name = subprocess.check_output(["c++filt", m.group(1)]).decode().split("(")[0]
        if flt in name:
            sm = re.search(r"(\d+) bytes smem", m.group(6))
            print("%-72s regs=%4s stack=%4s spill=%s/%s smem=%s" % (name, m.group(5), m.group(2), m.group(3), m.group(4), sm.group(1) if sm else 0))
