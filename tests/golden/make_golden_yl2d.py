#!/usr/bin/env python
"""Generate the Young-Laplace fixtures from the UNTOUCHED reference header AB/apps/Young_Laplace2D.h
(oracle/_ref/ref_yl2d; run where /root/reference exists):   make -C oracle ref && python tests/golden/make_golden_yl2d.py"""
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EXE = os.path.join(ROOT, "oracle", "_ref", "ref_yl2d")

CASES = {
    "yl2d_32x32_s200": dict(nx=32, ny=32, steps=200),                                             # config defaults
    "yl2d_32x32_s1000_long": dict(nx=32, ny=32, steps=1000),                                      # north_star horizon (CPU suite only)
    "yl2d_40x24_s150": dict(nx=40, ny=24, steps=150, Sigma=0.02, W=3.0, M=0.05, RhoL=0.01, RhoH=1.0, tau=0.7),
}


def run_case(name):
    kw = CASES[name]
    ne = kw["nx"] * kw["ny"]
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "d.bin")
        print(name, subprocess.check_output([EXE] + ["%s=%r" % kv for kv in kw.items()] + ["out=" + out]).decode().strip())
        raw = np.fromfile(out, dtype=np.uint8)
    data = {"params": np.frombuffer(json.dumps(kw).encode(), dtype=np.uint8)}
    off = 0
    for nm in ("C", "P", "Rho", "Ux", "Uy"):
        data[nm] = raw[off:off + 8 * ne].view(np.float64).copy()
        off += 8 * ne
    data["lattice"] = raw[off:off + 8 * 36 * ne].view(np.float64).copy()
    off += 8 * 36 * ne
    data["parity"] = np.array(int(raw[off:off + 4].view(np.int32)[0]))
    assert off + 4 == raw.size
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)


if __name__ == "__main__":
    import sys
    for n in (sys.argv[1:] or CASES):
        run_case(n)
