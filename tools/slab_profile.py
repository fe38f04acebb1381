#!/usr/bin/env python
"""Per-launch device time of ONE slab step on a self ring (CLBM_FORCE_SLAB=1, clbm_profile_step: events around every launch,
serialised) next to the time of a replayed step: what the small kernels of the ring protocol cost by themselves.
usage: python tools/slab_profile.py [sc3d|hcz3d|hcz2d] [nx]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
P, clbm = pkg.params, pkg.clbm
import bench  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "sc3d"
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 64
size = {"sc3d": (nx, 512, 512), "hcz3d": (nx, 512, 512), "hcz2d": (nx, 8194, 1)}[key]
os.environ["CLBM_FORCE_SLAB"] = "1"
prm, case, args = bench.build_params(P, key, *size, size[0], 0, 1)
if key == "sc3d":
    args = (0.265, 0.038, 0.2 * size[1], 5.0)
lat = clbm.Lattice(prm)
lat.init_case(case, args)
lat.peer_connect_local(lat, lat)
lat.slab_step(6)
lat.sync()
for rep in range(2):
    rows = lat.profile_step()
    print("%s %dx%dx%d step %d: " % ((key,) + size + (rep,)) + ", ".join("%s %.1f us" % (n, ms * 1e3) for n, ms in rows)
          + "  | sum %.1f us" % (1e3 * sum(ms for _, ms in rows)), flush=True)
lat.close()
