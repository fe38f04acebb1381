#!/usr/bin/env python
"""Generate the Pulsatile fixtures from the UNTOUCHED reference (run where /root/reference exists):

    make -C oracle ref && python tests/golden/make_golden_pulsatile.py

* pulsatile_N*_*.npz : binary dumps (fp64 P, Ux, Uy, yr1, yr2, u8 flag, both lattice buffers, parity) written by
  oracle/_ref/ref_pulsatile (the reference header AB/apps/PulsatileBloodFlow2D.h compiled where it lies, N a parameter).
* pulsatile_vtk_sha256.json : SHA-256 of the 103 sol_*.vtk files the reference ships in
  "Abbashub LBM/out_single-phase fluid flow through a compliant vessel/" (N = 64, the reference's own golden output).
"""
import glob
import hashlib
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EXE = os.path.join(ROOT, "oracle", "_ref", "ref_pulsatile")
VTK_DIR = "/root/reference/Abbashub LBM/out_single-phase fluid flow through a compliant vessel"

# name -> (N, dump steps, extra harness arguments)
CASES = {
    "pulsatile_N16_s1_50_600": (16, [1, 50, 600], {}),                       # deformable, severed (driver defaults)
    "pulsatile_N24_s400": (24, [400], {}),                                    # many fresh nodes
    "pulsatile_N20_rigid_s300": (20, [300], {"deformable": 0}),               # rigid walls
    "pulsatile_N32_tau08_s1500": (32, [1500], {"tau": 0.8}),
}


def run_case(name):
    N, dumps, kw = CASES[name]
    nx, ny = 1 + 10 * (N - 2), N
    ne = nx * ny
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "d.bin")
        args = [EXE, "N=%d" % N, "steps=%d" % max(dumps), "out=" + out, "dump_at=" + ",".join(map(str, dumps))]
        args += ["%s=%r" % kv for kv in kw.items()]
        print(name, subprocess.check_output(args).decode().strip())
        raw = np.fromfile(out, dtype=np.uint8)
    data = {"params": np.frombuffer(json.dumps(dict(N=N, dumps=dumps, kw=kw)).encode(), dtype=np.uint8)}
    off = 0
    for d in dumps:
        for nm, n in (("P", ne), ("Ux", ne), ("Uy", ne), ("yr1", nx), ("yr2", nx)):
            data["%s_%d" % (nm, d)] = raw[off:off + 8 * n].view(np.float64).copy()
            off += 8 * n
        data["flag_%d" % d] = raw[off:off + ne].copy()
        off += ne
        lat = raw[off:off + 8 * 18 * ne].view(np.float64).copy()
        off += 8 * 18 * ne
        par = int(raw[off:off + 4].view(np.int32)[0])
        off += 4
        data["parity_%d" % d] = np.array(par)
        # keep the full lattice only for the last dump (size); a checksum-free exact copy of both buffers
        if d == dumps[-1]:
            data["lattice_%d" % d] = lat
    assert off == raw.size
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)


def vtk_hashes():
    h = {}
    for f in sorted(glob.glob(os.path.join(VTK_DIR, "sol_*.vtk"))):
        h[os.path.basename(f)] = hashlib.sha256(open(f, "rb").read()).hexdigest()
    assert len(h) == 103
    json.dump(h, open(os.path.join(HERE, "pulsatile_vtk_sha256.json"), "w"), indent=0, sort_keys=True)
    print("hashed", len(h), "reference VTK files")


if __name__ == "__main__":
    for n in CASES:
        run_case(n)
    vtk_hashes()
