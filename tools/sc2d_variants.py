"""Tile / stage shapes of the D2Q9 Shan-Chen TMA kernel (sc2d_tma.cu) at 8192 x 8192 against the register-pipelined kernel.
   python tools/sc2d_variants.py [steps]      (ctypes only: no torch import)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
P = pkg.params
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
os.environ["CLBM_SC_MULTI"] = "0"
for tma in (0, 1, 5, 7, 8, 9, 11, 12, 13, 14, 15):
    out = []
    for xc in (None, 64):
        os.environ["CLBM_SC2D_TMA"] = str(tma)
        if xc:
            os.environ["CLBM_SC_XCHUNK"] = str(xc)
        else:
            os.environ.pop("CLBM_SC_XCHUNK", None)
        prm = P.sc_params(P.MODEL_SC_D2Q9, 8192, 8192, tau=1.0)
        with pkg.clbm.Lattice(prm) as lat:
            lat.init_case(P.CASE_SC_LAPLACE2D, (0.265, 0.038, 10.0))
            lat.step(5)
            lat.sync()
            ms = lat.step_timed(steps) / steps
        out.append("xchunk %s: %.3f ms %.0f MLUPS" % (xc or "dflt", ms, 8192 * 8192 / ms / 1e3))
    print("CLBM_SC2D_TMA=%d (%s): %s" % (tma, "register-pipelined kernel" if tma == 0 else "TMA kernel", ", ".join(out)), flush=True)
