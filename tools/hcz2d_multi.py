"""BASELINE configs[1] (HCZ D2Q9 256 x 1026, L2 resident): launch-per-step against the multi-step cooperative launch (hcz2d_fused.cu,
MULTI form), over x-chunk lengths; populations must be bit-identical.
   python tools/hcz2d_multi.py [steps]      (ctypes only: no torch import)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
P = pkg.params
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000


def run(prm, case, args, multi, xchunk):
    os.environ["CLBM_HCZ2D_MULTI"] = str(multi)
    if xchunk:
        os.environ["CLBM_HCZ2D_XCHUNK"] = str(xchunk)
    else:
        os.environ.pop("CLBM_HCZ2D_XCHUNK", None)
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
        ("c2 HCZ D2Q9 256x1026", P.hcz_params(P.MODEL_HCZ_D2Q9, 256, 1026, N=256), P.CASE_HCZ_RT2D, ()),
        ("HCZ D2Q9 MRT 256x1026", P.hcz_mrt_params(256, 1026, N=256), P.CASE_HCZ_RT2D, ()),
        ("HCZ D2Q9 128x514", P.hcz_params(P.MODEL_HCZ_D2Q9, 128, 514, N=128), P.CASE_HCZ_RT2D, ())):
    base_us, base, nl = run(prm, case, args, 0, None)
    out = ["launch per step %.2f us (%d launches)" % (base_us, nl)]
    for xc in (None, 6, 8, 16):
        us, pops, nl = run(prm, case, args, 1, xc)
        out.append("multi xchunk %s: %.2f us%s (%d launches)" % (xc or "auto", us, "" if np.array_equal(pops, base) else " MISMATCH", nl))
    print("%s, %d steps/run, %.0f nodes: %s" % (name, steps, prm.nelem, "; ".join(out)), flush=True)
