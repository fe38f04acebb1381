"""L2-resident lattices (BASELINE configs[0], configs[1]) are latency bound: a CTA marches its x-chunk serially.  This sweeps
the x-chunk length of the fused D2Q9 kernels on one GPU, checks that every chunking gives bit-identical populations, and
prints microseconds per step.   python tools/small_lattice_chunks.py [steps]      (ctypes only: no torch import)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
P = pkg.params
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
t_start = time.time()


def run(prm, case, args, env, val):
    if val is None:
        os.environ.pop(env, None)
    else:
        os.environ[env] = str(val)
    with pkg.clbm.Lattice(prm) as lat:
        lat.init_case(case, args)
        lat.step(20)
        lat.sync()
        ms = lat.step_timed(steps)
        pops = lat.in_pops()
    os.environ.pop(env, None)
    return ms * 1e3 / steps, pops


for name, prm, case, args, env in (
        ("c1 SC D2Q9 256x256", P.sc_params(P.MODEL_SC_D2Q9, 256, 256, ulb=0.01, N=256, Re=6.0), P.CASE_SC_LAPLACE2D, (0.265, 0.038, 10.0), "CLBM_SC_XCHUNK"),
        ("c2 HCZ D2Q9 256x1026", P.hcz_params(P.MODEL_HCZ_D2Q9, 256, 1026, N=256), P.CASE_HCZ_RT2D, (), "CLBM_HCZ2D_XCHUNK")):
    base_us, base = run(prm, case, args, env, None)
    out = ["default %.1f us" % base_us]
    for xc in (8, 4, 2, 1):
        us, pops = run(prm, case, args, env, xc)
        out.append("xchunk %d: %.1f us%s" % (xc, us, "" if np.array_equal(pops, base) else " MISMATCH"))
    us, pops = run(prm.copy(fused=0), case, args, env, None)
    out.append("staged: %.1f us (max rel diff vs fused %.1e)" % (us, np.max(np.abs(pops - base)) / np.max(np.abs(base))))
    print("%s, %d steps/run, MLUPS = %.0f / us: %s" % (name, steps, prm.nelem, ", ".join(out)), flush=True)
print("wall %.1f s" % (time.time() - t_start))
