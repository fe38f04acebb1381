"""error growth GPU vs oracle for candidate Young-Laplace parameter sets (debug tool)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, _cases
from _oracle import YL2DOracle
clbm = _cases.pkg.clbm
cands = [((128, 128), {}, 1000), ((64, 48), dict(Sigma=0.005, W=4.0, M=0.02, RhoL=0.1, RhoH=1.0, tau=0.8), 600),
         ((96, 64), dict(Sigma=0.02, W=5.0, M=0.05, RhoL=0.01, tau=0.7), 1000), ((64, 64), dict(Sigma=0.01, W=5.0, M=0.03, RhoL=0.05, tau=0.9), 1000),
         ((48, 40), {}, 200)]
for (nx, ny), kw, steps in cands:
    o = YL2DOracle(nx, ny, **kw); d = clbm.YoungLaplace(nx, ny, **kw)
    done = 0
    for chunk in (steps // 4,) * 4:
        o.step(chunk); d.step(chunk); done += chunk
        fo, fd = o.fields(), d.fields()
        print((nx, ny), kw, done, {k: "%.1e" % _cases.rel_linf(fd[k], fo[k]) for k in fo}, "Umax %.2e" % np.abs(fo["Ux"]).max(), flush=True)
    o.close(); d.close()
