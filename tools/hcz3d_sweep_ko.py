"""Knock-out timing of the HCZ D3Q19 single-sweep kernel (hcz3d_sweep.cu, CLBM_HCZ3D_SWEEP_KO): which phase is on the critical
path?  Results of a knocked-out run are WRONG by construction; only the time per step is read.
   python tools/hcz3d_sweep_ko.py [nx] [steps] [ko ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
P = pkg.params
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
kos = [int(v) for v in sys.argv[3:]] or [0, 1, 2, 4, 8, 16, 32, 64, 128, 256]
NAMES = {1: "no ring gather (S5)", 2: "no gather at all (S5)", 4: "no ring level 2 (S3)", 8: "no ring moment loads (S1)",
         16: "no thread-level population stores (S4)", 1024: "no TMA box stores", 32: "no S2", 64: "no edge loads", 128: "no collide (S4)", 256: "no S3"}
for ko in kos:
    os.environ["CLBM_HCZ3D_SWEEP_KO"] = str(ko)
    prm = P.hcz_params(P.MODEL_HCZ_D3Q19, nx, 512, 512, ulb=0.01, N=nx, Re=6.0, kappa=5e-4, gravity=0.0)
    prm.nx_global, prm.x_offset, prm.fused = nx, 0, 1
    with pkg.clbm.Lattice(prm) as lat:
        lat.init_case(P.CASE_HCZ_LAPLACE3D, ())
        lat.step(3)
        lat.sync()
        ms = min(lat.step_timed(steps) / steps for _ in range(2))
    what = " + ".join(NAMES[b] for b in sorted(NAMES) if ko & b) or "full kernel"
    print("ko %4d: %7.3f ms per step   (%s)" % (ko, ms, what), flush=True)
