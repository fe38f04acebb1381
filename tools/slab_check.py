#!/usr/bin/env python
"""Multi-GPU slab check (run under torchrun, one rank per GPU):
every rank advances its x-slab over NCCL (DistRing); rank 0 also advances the whole lattice on its own GPU and
compares the gathered populations bit-for-bit.   torchrun --nproc-per-node N tools/slab_check.py [--native]
--transport=peer|nccl|torch: the ghost exchange (default peer: the library's peer-memory ring over CUDA IPC with CUDA-graph
replay; nccl = --native: library-driven ncclSend/ncclRecv; torch: batch_isend_irecv issued from Python)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
P, clbm, slab = pkg.params, pkg.clbm, pkg.slab


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    cases = [
        ("sc3d", P.sc_params(P.MODEL_SC_D3Q19, 8 * world + 4, 16, 24, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT),
         P.CASE_SC_DROPLET3D, (0.265, 0.038, 6.0, 5.0), 40),
        ("hcz2d", P.hcz_params(P.MODEL_HCZ_D2Q9, 8 * world, 66, N=8 * world), P.CASE_HCZ_RT2D, (), 40),
        ("hcz3d", P.hcz_params(P.MODEL_HCZ_D3Q19, 6 * world, 12, 12, ulb=0.01, N=6 * world, Re=6.0, kappa=5e-4, gravity=-1e-5),
         P.CASE_HCZ_LAPLACE3D, (), 25),
    ]
    ok_all = True
    for name, prm, case, args, steps in cases:
        sp = slab.slab_params(prm, rank, world)
        sp.device = lr
        lat = clbm.Lattice(sp)
        lat.init_case(case, args)
        transport = None
        for a in sys.argv[1:]:
            if a.startswith("--transport="):
                transport = a.split("=", 1)[1]
        ring = slab.DistRing(lat, rank, world, dev, native="--native" in sys.argv, transport=transport)
        ring.step(steps)
        pops = torch.from_numpy(lat.in_pops()).to(dev)          # [sets, Q, nelem_local]
        sizes = [b[1] - b[0] for b in slab.slab_bounds(prm.nx_global, world)]
        plane = prm.ny * prm.nz
        parts = [torch.empty((prm.sets, prm.Q, s * plane), dtype=torch.float64, device=dev) for s in sizes]
        dist.all_gather(parts, pops)
        lat.close()
        if rank == 0:
            full = torch.cat(parts, dim=2).cpu().numpy()
            single = clbm.Lattice(prm.copy(device=lr))    # same kernel variant as the slabs
            single.init_case(case, args)
            single.step(steps)
            ref = single.in_pops()
            single.close()
            same = np.array_equal(full, ref)
            err = np.max(np.abs(full - ref)) / np.max(np.abs(ref))
            print("%-6s world=%d  transport=%s  bit-identical=%s  rel Linf=%.3e" % (name, world, ring.transport, same, err), flush=True)
            ok_all = ok_all and same
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok_all:
        sys.exit(1)


if __name__ == "__main__":
    main()
