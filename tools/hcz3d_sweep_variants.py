"""Load-issue variants of the HCZ D3Q19 single-sweep kernel (hcz3d_sweep.cu, CLBM_HCZ3D_SWEEP_VAR) on the bench lattice.
Every variant is first checked against variant 0 (populations of a 16x512x512 sub-run after 7 steps, bit for bit), then timed.
   python tools/hcz3d_sweep_variants.py [nx] [steps] [variant ...]      (ctypes only: no torch import)"""
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
variants = [int(v) for v in sys.argv[3:]] or [0, 1, 2, 3]
t_start = time.time()


def lattice(n, variant):
    os.environ["CLBM_HCZ3D_SWEEP_VAR"] = str(variant)      # read once per context, in clbm_create
    prm = P.hcz_params(P.MODEL_HCZ_D3Q19, n, 512, 512, ulb=0.01, N=n, Re=6.0, kappa=5e-4, gravity=0.0)
    prm.nx_global, prm.x_offset, prm.fused = n, 0, 1
    lat = pkg.clbm.Lattice(prm)
    lat.init_case(P.CASE_HCZ_LAPLACE3D, ())
    return lat


def check(variant):
    with lattice(16, variant) as lat:
        lat.step(7)
        lat.sync()
        return lat.in_pops()


base = check(0)
for v in variants:
    got = check(v)
    ok = np.array_equal(got, base)
    diff = float(np.max(np.abs(got - base)) / np.max(np.abs(base)))
    del got
    with lattice(nx, v) as lat:
        lat.step(5)
        lat.sync()
        best = min(lat.step_timed(steps) / steps for _ in range(3))
    print("variant %d (%s): %.3f ms per step, %.0f MLUPS, %.3f of 6545 GB/s" %
          (v, "bit-identical to 0" if ok else "max rel diff vs 0 %.1e" % diff, best, nx * 512 * 512 / best / 1e3,
           609 * nx * 512 * 512 / best / 1e6 / 6545.0), flush=True)
print("wall %.1f s" % (time.time() - t_start))
