"""BASELINE configs[0] (Shan-Chen D2Q9 256 x 256, L2 resident): launch-by-launch steps against the multi-step cooperative launch
(sc_fused.cu, MULTI form: grid barrier between steps), over tile heights and x-chunk lengths; populations must be bit-identical.
   python tools/small_lattice_multi.py [steps]      (ctypes only: no torch import)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
P = pkg.params
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000


def run(prm, case, args, multi, xchunk):
    os.environ["CLBM_SC_MULTI"] = str(multi)
    if xchunk:
        os.environ["CLBM_SC_XCHUNK"] = str(xchunk)
    else:
        os.environ.pop("CLBM_SC_XCHUNK", None)
    with pkg.clbm.Lattice(prm) as lat:
        lat.init_case(case, args)
        lat.step(20)
        lat.sync()
        ms = lat.step_timed(steps)
        lat.step(1)          # an odd total: the parity bookkeeping of the multi-step launch is part of the check
        pops = lat.in_pops()
        launches = lat.launch_count()
    return ms * 1e3 / steps, pops, launches


for name, prm, case, args in (
        ("c1 SC D2Q9 256x256", P.sc_params(P.MODEL_SC_D2Q9, 256, 256, ulb=0.01, N=256, Re=6.0), P.CASE_SC_LAPLACE2D, (0.265, 0.038, 10.0)),
        ("SC D2Q9 gravity 200x130", P.sc_params(P.MODEL_SC_D2Q9, 200, 130, omega=1.2, gravity=-1e-5), P.CASE_SC_LAPLACE2D, (0.265, 0.038, 20.0)),
        ("SC D2Q9 contact 512x256 walls", P.sc_params(P.MODEL_SC_D2Q9, 512, 256, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT), P.CASE_SC_CONTACT2D, (0.265, 0.038, 40.0))):
    base_us, base, nl = run(prm, case, args, 0, None)
    out = ["launch by launch %.2f us (%d launches)" % (base_us, nl)]
    os.environ.pop("CLBM_SC_MULTI", None)
    us, pops, nl = run(prm, case, args, -1, None)
    out.append("default (column-resident kernel with a grid barrier where the columns fit, else plane marching): %.2f us%s (%d launches)" % (us, "" if np.array_equal(pops, base) else " MISMATCH", nl))
    us, pops, nl = run(prm, case, args, 6, None)
    out.append("column-resident kernel, neighbour flags: %.2f us%s (%d launches)" % (us, "" if np.array_equal(pops, base) else " MISMATCH", nl))
    for multi, label in ((1, "128-row CTAs"), (2, "256-row CTAs")):
        for xc in (1, 2):
            us, pops, nl = run(prm, case, args, multi, xc)
            out.append("multi %s xchunk %d: %.2f us%s (%d launches)" % (label, xc, us, "" if np.array_equal(pops, base) else " MISMATCH", nl))
    print("%s, %d steps/run, %.0f nodes: %s" % (name, steps, prm.nelem, "; ".join(out)), flush=True)
