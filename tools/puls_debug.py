"""find the first step at which the device pulsatile path departs from the oracle (debug tool)"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import _cases
from _oracle import PulsatileOracle
clbm = _cases.pkg.clbm
N = int(sys.argv[1]); steps = int(sys.argv[2]); kw = eval(sys.argv[3]) if len(sys.argv) > 3 else {}
o = PulsatileOracle(N=N, **kw); d = clbm.Pulsatile(N=N, **kw)
ny = o.ny
for s in range(steps):
    o.step(1); d.step(1)
    fo, fd = o.fields(), d.fields()
    lo, ld = o.lattice(), d.lattice()[0]
    bad = [k for k in fo if not np.array_equal(fo[k], fd[k])]
    if bad or not np.array_equal(lo, ld):
        print("first mismatch after step", s + 1, bad, "parity", o.parity)
        for k in bad:
            idx = np.nonzero(fo[k] != fd[k])[0]
            print(k, len(idx), [(int(i // ny), int(i % ny)) if fo[k].size == o.nelem else int(i) for i in idx[:10]], fo[k][idx[:4]], fd[k][idx[:4]])
        idx = np.nonzero(lo != ld)[0]
        ne = o.nelem
        print("lattice", len(idx), [(int(i // (9 * ne)), int((i % (9 * ne)) // ne), int((i % ne) // ny), int(i % ny)) for i in idx[:20]])
        print(lo[idx[:6]], ld[idx[:6]])
        print("yr1[:4]", fo["yr1"][:4], "yr2[:4]", fo["yr2"][:4])
        X = int((idx[0] % ne) // ny) if len(idx) else 0
        print("at X", X, "yr1", fo["yr1"][max(0, X-2):X+3], "yr2", fo["yr2"][max(0, X-2):X+3])
        print("flag col", fo["flag"][X*ny:(X+1)*ny])
        break
else:
    print("all", steps, "steps identical")
