"""Tile / stage-count / release-order variants of the D3Q19 Shan-Chen TMA kernel (sc_fused_tma.cu) on the bench lattice.
Every variant is checked against the default one (populations of a 64x512x512 sub-run bit-identical) before it is timed.
   python tools/sc3d_variants.py [nx] [steps] [variant ...]      (ctypes only: no torch import)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
P = pkg.params
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
variants = [int(v) for v in sys.argv[3:]] or [11, 29, 41]
t_start = time.time()


def lattice(n, variant, xchunk=None):
    os.environ["CLBM_SC_TILE"] = str(variant)
    if xchunk:
        os.environ["CLBM_SC_XCHUNK"] = str(xchunk)
    else:
        os.environ.pop("CLBM_SC_XCHUNK", None)
    prm = P.sc_params(P.MODEL_SC_D3Q19, n, 512, 512, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT)
    lat = pkg.clbm.Lattice(prm)
    lat.init_case(P.CASE_SC_DROPLET3D, (0.265, 0.038, 0.2 * 512, 5.0))
    return lat


def check(variant):
    with lattice(16, variant) as lat:
        lat.step(7)
        lat.sync()
        return lat.in_pops()


base = check(11)
for v in variants:
    got = check(v)
    ok = np.array_equal(got, base)
    diff = float(np.max(np.abs(got - base)) / np.max(np.abs(base)))
    out = []
    for xc in (None, 16, 32, 48):
        with lattice(nx, v, xc) as lat:
            lat.step(5)
            lat.sync()
            ms = lat.step_timed(steps) / steps
        out.append("xchunk %s: %.3f ms %.0f MLUPS" % (xc or "dflt", ms, nx * 512 * 512 / ms / 1e3))
    print("variant %d (%s): %s" % (v, "bit-identical to 11" if ok else "max rel diff vs 11 %.1e" % diff, ", ".join(out)), flush=True)
print("wall %.1f s" % (time.time() - t_start))
