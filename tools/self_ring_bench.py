#!/usr/bin/env python
"""Slab-step overhead on ONE GPU: a full-width lattice forced into x-slab mode (CLBM_FORCE_SLAB=1) on a peer ring with itself.
Same stages, packs, signal / wait kernels, overlap protocol and graph replay as a rank of a multi-GPU run (minus the NVLink
hop), so the difference to the plain single-slab step of the same lattice is the protocol's own cost at that slab size.
usage: python tools/self_ring_bench.py [sc3d|hcz3d|hcz2d|sc2d] [nx] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
P, clbm = pkg.params, pkg.clbm
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def run(key, size, steps, force_slab, graph, overlap=-1):
    import torch
    os.environ["CLBM_FORCE_SLAB"] = "1" if force_slab else "0"
    os.environ["CLBM_SLAB_GRAPH"] = str(graph)
    if overlap >= 0:
        os.environ["CLBM_SLAB_OVERLAP"] = str(overlap)
    else:
        os.environ.pop("CLBM_SLAB_OVERLAP", None)
    prm, case, args = bench.build_params(P, key, *size, size[0], 0, 1)
    if key == "sc3d":
        args = (0.265, 0.038, 0.2 * size[1], 5.0)
    lat = clbm.Lattice(prm)
    lat.init_case(case, args)
    if force_slab:
        lat.peer_connect_local(lat, lat)
        step = lat.slab_step
    else:
        step = lat.step
    step(5); step(3); step(4)
    lat.sync()
    stream = torch.cuda.ExternalStream(lat.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lat.launch_count()
    e0.record(stream)
    step(steps)
    e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = (lat.launch_count() - l0) / steps
    lat.kernel_timing_begin(8)
    step(5)
    lat.sync()
    kms, kcount, kname = lat.kernel_timing_end()
    lat.close()
    return ms, launches, kms, kname


def main():
    key = sys.argv[1] if len(sys.argv) > 1 else "sc3d"
    nx = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    size = {"sc3d": (nx, 512, 512), "hcz3d": (nx, 512, 512), "hcz2d": (nx, 8194, 1), "sc2d": (nx, 8192, 1)}[key]
    nelem = size[0] * size[1] * size[2]
    for name, fs, gr, ov in (("single slab (clbm_step)", 0, 0, -1), ("self ring, sequential, eager", 1, 0, 0),
                             ("self ring, sequential, graph", 1, 1, 0), ("self ring, interior first, graph", 1, 1, 1),
                             ("self ring, halo first, graph", 1, 1, 2)):
        ms, launches, kms, kname = run(key, size, steps, fs, gr, ov)
        print("%-5s %dx%dx%d  %-34s %8.1f us/step  %7.0f MLUPS  %5.1f launches/step  dominant kernel %s %.1f us"
              % (key, size[0], size[1], size[2], name, ms * 1e3, nelem / ms / 1e3, launches, kname, kms * 1e3), flush=True)


if __name__ == "__main__":
    main()
