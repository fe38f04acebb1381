"""Host-side slab logic on CPU: decomposition arithmetic, host-array slicing, and the ring exchange
pattern over torch.distributed with the gloo backend at world_size 2 and 3 (no GPU needed)."""
import os
import socket

import numpy as np
import pytest

import _cases

pkg = _cases.pkg
P = pkg.params
slab = pkg.slab


def test_slab_bounds_cover_the_lattice():
    for nx, R in [(256, 8), (10, 3), (7, 7), (2048, 8), (513, 4)]:
        b = slab.slab_bounds(nx, R)
        assert b[0][0] == 0 and b[-1][1] == nx
        assert all(b[i][1] == b[i + 1][0] for i in range(R - 1))
        w = [x1 - x0 for x0, x1 in b]
        assert max(w) - min(w) <= 1


def test_slice_host_state_matches_reference_layout():
    prm = P.hcz_params(P.MODEL_HCZ_D3Q19, 6, 3, 4)
    ne = prm.nelem
    lattice = np.arange(prm.lattice_size, dtype=np.float64)
    flag = (np.arange(ne) % 2).astype(np.uint8)
    parts = [slab.slice_host_state(prm, lattice, flag, r, 3) for r in range(3)]
    plane = 12
    full = lattice.reshape(2, 2, 19, ne)
    for r, (l, f) in enumerate(parts):
        sp = slab.slab_params(prm, r, 3)
        assert sp.nx == 2 and sp.x_offset == 2 * r and sp.nx_global == 6
        np.testing.assert_array_equal(l.reshape(2, 2, 19, 2 * plane), full[..., 2 * r * plane:(2 * r + 2) * plane])
        np.testing.assert_array_equal(f, flag[2 * r * plane:(2 * r + 2) * plane])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ring_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        send0 = torch.full((5,), 10.0 * rank + 0)   # towards rank-1
        send1 = torch.full((5,), 10.0 * rank + 1)   # towards rank+1
        recv0 = torch.zeros(5)
        recv1 = torch.zeros(5)
        for _ in range(3):
            slab.ring_exchange(dist, rank, world, send0, send1, recv0, recv1)
        left, right = (rank - 1) % world, (rank + 1) % world
        ok = bool((recv0 == 10.0 * left + 1).all() and (recv1 == 10.0 * right + 0).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_exchange_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ring_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


# ---------------------------------------------------------------------------------------------------------------------------
# Slab-aware diagnostics: DistRing.contact_angle_scan / interface_heights combine the per-slab integers of the device scans
# (csrc/diag_kernels.cu:43-50) with MIN / MAX all-reduces.  Here every rank holds a numpy stand-in for its slab's device scan
# (same contract, global x) and the ring's result must equal the reference's SERIAL scan of the whole lattice
# (SC/apps/contactAngle2D.h:465-505, restated below) -- over gloo at world_size 2 and 3, uneven slabs included.
# ---------------------------------------------------------------------------------------------------------------------------
def _serial_contact_angle(rho, flag, rho_cut):
    """calculateContactAngle's integers (base_y, Base, Height) on the whole lattice, rho/flag as [nx][ny]"""
    nx, ny = rho.shape
    base_y = 2
    while base_y < ny and flag[0, base_y] == 0:
        base_y += 1
    if base_y >= ny - 1:
        return base_y, 0, 0
    xmid = nx // 2
    left = right = xmid
    while left > 0 and rho[(left - 1) % nx, base_y] > rho_cut:
        left -= 1
    while right < nx - 1 and rho[(right + 1) % nx, base_y] > rho_cut:
        right += 1
    base = max(0, right - left + 1)
    height = 0
    for y in range(base_y, ny):
        if flag[xmid, y] == 0 or not rho[xmid, y] > rho_cut:
            break
        height += 1
    return base_y, base, height


class _SlabScanStandIn:
    """what clbm_diag_contact_angle_slab returns for the columns [x0, x1) of the global lattice"""

    def __init__(self, rho, flag, x0, x1):
        self.rho, self.flag, self.x0, self.x1 = rho, flag, x0, x1
        self.p = type("p", (), {"ny": rho.shape[1]})()

    def contact_angle_scan_slab(self, rho_cut, base_y_in):
        nxg, ny = self.rho.shape
        xmid = nxg // 2
        base_y = base_y_in
        if base_y_in < 0:
            base_y = ny
            if self.x0 == 0:
                base_y = 2
                while base_y < ny and self.flag[0, base_y] == 0:
                    base_y += 1
            return [base_y, -1, nxg, ny]
        lstop, rstop, hstop = -1, nxg, ny
        for x in range(self.x0, self.x1):
            if not self.rho[x, base_y] > rho_cut:
                if x < xmid:
                    lstop = max(lstop, x)
                elif x > xmid:
                    rstop = min(rstop, x)
        if self.x0 <= xmid < self.x1:
            hstop = next((y for y in range(base_y, ny) if self.flag[xmid, y] == 0 or not self.rho[xmid, y] > rho_cut), ny)
        return [base_y, lstop, rstop, hstop]


def _serial_interface_heights(phi, phi_mid):
    """findInterfaceHeights (PF/apps/rayleighTaylor2D.h:668-708): walking down from y = ny-2 to 1 on the columns x = 0 and
    x = nx/2, the first y with phi <= phi_mid; both outputs start from int(+-0.05) = 0"""
    nx, ny = phi.shape
    out = []
    for x in (0, nx // 2):
        out.append(next((y for y in range(ny - 2, 0, -1) if phi[x, y] <= phi_mid), 0))
    return tuple(out)


class _SlabHeightsStandIn:
    """what clbm_diag_interface_heights returns on the columns [x0, x1): 0 for a column this slab does not own"""

    def __init__(self, phi, x0, x1):
        self.phi, self.x0, self.x1 = phi, x0, x1

    def interface_heights(self, phi_mid):
        nxg, ny = self.phi.shape
        out = []
        for x in (0, nxg // 2):
            c = 0
            if self.x0 <= x < self.x1:
                for y in range(1, ny - 1):
                    if self.phi[x, y] <= phi_mid:
                        c = y
            out.append(c)
        return out


def _rt_interface(nx, ny, amp):
    import numpy as np
    x, y = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    h = ny / 2 + amp * nx * np.cos(2 * np.pi * x / (nx - 1))
    return 0.1375 + 0.1135 * np.tanh((y - h) / 1.25)       # heavy (phi_l) above, light below


def _sessile_droplet(nx, ny, radius, wall_rows):
    import numpy as np
    x, y = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    r = np.sqrt((x - nx / 2 + 0.3) ** 2 + (y - wall_rows) ** 2)
    rho = 0.1515 - 0.1135 * np.tanh((r - radius) / 1.5)
    flag = np.ones((nx, ny), dtype=np.uint8)
    flag[:, :wall_rows] = 0
    flag[:, ny - 1] = 0
    return rho, flag


def _diag_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = []
        for nx, ny, radius, wall_rows in ((37, 24, 9.0, 1), (40, 30, 12.5, 3), (16, 12, 40.0, 1), (12, 6, 3.0, 5)):
            rho, flag = _sessile_droplet(nx, ny, radius, wall_rows)
            x0, x1 = slab.slab_bounds(nx, world)[rank]
            ring = object.__new__(slab.DistRing)          # the reduction logic only: no device, no halo buffers
            ring.torch, ring.dist, ring.R, ring.rank, ring.dev = torch, dist, world, rank, torch.device("cpu")
            ring.lat = _SlabScanStandIn(rho, flag, x0, x1)
            out.append(tuple(ring.contact_angle_scan(0.1515)))
        heights = []
        for nx, ny, amp in ((24, 98, 0.1), (33, 70, 0.25), (16, 40, 0.0)):
            phi = _rt_interface(nx, ny, amp)
            ring = object.__new__(slab.DistRing)
            ring.torch, ring.dist, ring.R, ring.rank, ring.dev = torch, dist, world, rank, torch.device("cpu")
            ring.lat = _SlabHeightsStandIn(phi, *slab.slab_bounds(nx, world)[rank])
            heights.append(tuple(ring.interface_heights(0.1375)))
        q.put((rank, (out, heights)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_contact_angle_reduction_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_diag_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    want = [_serial_contact_angle(*_sessile_droplet(nx, ny, radius, wall_rows), 0.1515)
            for nx, ny, radius, wall_rows in ((37, 24, 9.0, 1), (40, 30, 12.5, 3), (16, 12, 40.0, 1), (12, 6, 3.0, 5))]
    assert want[0][1] > 5 and want[0][2] > 3            # a droplet is found ...
    assert want[2][1] == 16                              # ... a liquid film spans the whole row (no stop on either side) ...
    assert want[3][1:] == (0, 0)                         # ... and a lattice without a fluid row above the wall
    want_h = [_serial_interface_heights(_rt_interface(nx, ny, amp), 0.1375) for nx, ny, amp in ((24, 98, 0.1), (33, 70, 0.25), (16, 40, 0.0))]
    assert want_h[0][0] != want_h[0][1] and min(want_h[0]) > 10      # the perturbed interface sits at different heights on the two columns
    for r in range(world):
        assert res[r][0] == want, (r, res[r][0], want)        # every rank holds the serial scan's integers
        assert res[r][1] == want_h, (r, res[r][1], want_h)    # spike / bubble heights (findInterfaceHeights)
    # the in-process combine used by LocalRing (same contract)
    for (nx, ny, radius, wall_rows), w in zip(((37, 24, 9.0, 1), (40, 30, 12.5, 3), (16, 12, 40.0, 1), (12, 6, 3.0, 5)), want):
        rho, flag = _sessile_droplet(nx, ny, radius, wall_rows)
        slabs = [_SlabScanStandIn(rho, flag, *b) for b in slab.slab_bounds(nx, world)]
        first = [s.contact_angle_scan_slab(0.1515, -1) for s in slabs]
        base_y = min(f[0] for f in first)
        second = [s.contact_angle_scan_slab(0.1515, base_y) for s in slabs] if base_y < ny - 1 else None
        assert slab.combine_contact_angle(ny, first, second) == w
