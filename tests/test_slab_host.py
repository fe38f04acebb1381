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
