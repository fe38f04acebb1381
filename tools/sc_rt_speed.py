"""MLUPS of the Shan-Chen Rayleigh-Taylor variant (fused / staged kernel) on one GPU: python tools/sc_rt_speed.py [N] [steps]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import _cases  # noqa: E402

pkg = _cases.pkg
P = pkg.params
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for fused in (1, 0):
    prm = P.sc_rt_params(N, omega=1.0).copy(fused=fused)
    with pkg.clbm.Lattice(prm) as lat:
        lat.init_case(P.CASE_SC_RT2D, (1.2, 0.4))
        lat.step(10)
        lat.sync()
        ms = lat.step_timed(steps)
        m = lat.reduce(P.REDUCE_MASS)
    print("sc_rt2d %dx%d fused=%d: %.1f MLUPS (%.3f ms/step, 145 B/LU -> %.0f GB/s), mass %.6f"
          % (prm.nx, prm.ny, fused, prm.nelem * steps / (ms * 1e3), ms / steps, 145 * prm.nelem * steps / (ms * 1e-3) / 1e9, m))
