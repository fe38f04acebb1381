#!/usr/bin/env python
"""Where does a slab step of the overlap protocol spend its time?  CUDA events on the launching stream and on the
boundary stream after every stage of a few steps (run under torchrun, one rank per GPU)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
P, clbm, slab = pkg.params, pkg.clbm, pkg.slab


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    nxl = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    prm = P.sc_params(P.MODEL_SC_D3Q19, nxl, 512, 512, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT)
    prm.nx_global, prm.x_offset, prm.device = nxl * world, nxl * rank, lr
    lat = clbm.Lattice(prm)
    lat.init_case(P.CASE_SC_DROPLET3D, (0.265, 0.038, 0.2 * 512, 5.0))
    ring = slab.DistRing(lat, rank, world, dev)
    ring.step(5)
    lat.sync()
    ev = lambda s: (lambda e: (e.record(s), e)[1])(torch.cuda.Event(enable_timing=True))   # noqa: E731
    rows = []
    for _ in range(4):
        t0 = ev(ring.stream)
        b0 = ev(ring.stream_b)
        lat.step_stage(10)
        m1 = ev(ring.stream)          # interior done
        b1 = ev(ring.stream_b)        # psi + pack0 done
        ring.exchange(0, boundary=True)
        b2 = ev(ring.stream_b)
        lat.step_stage(11)
        b3 = ev(ring.stream_b)        # unpack0 + boundary collide + pack1 done
        ring.exchange(1, boundary=True)
        b4 = ev(ring.stream_b)
        lat.step_stage(12)
        b5 = ev(ring.stream_b)
        m2 = ev(ring.stream)          # joined
        rows.append((t0, b0, m1, b1, b2, b3, b4, b5, m2))
    torch.cuda.synchronize()
    if rank == 0:
        for t0, b0, m1, b1, b2, b3, b4, b5, m2 in rows:
            f = lambda e: "%7.3f" % t0.elapsed_time(e)   # noqa: E731
            print("interior done", f(m1), "| b: start", f(b0), "psi+pack0", f(b1), "xchg0", f(b2), "bnd collide+pack1", f(b3),
                  "xchg1", f(b4), "unpack1", f(b5), "| joined", f(m2), flush=True)
    lat.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
