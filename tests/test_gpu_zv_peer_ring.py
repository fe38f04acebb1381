"""The library-driven peer-memory ring (csrc/slab_comm.cu) on ONE GPU: R slab contexts of this process are connected with
clbm_peer_connect_local, so every pack writes straight into the neighbour's mailbox, the exchanges are the flag signal / wait
kernels, and clbm_slab_step replays two captured steps per CUDA-graph launch -- exactly the code path of a multi-GPU run,
minus the IPC mapping.  Results must equal the single slab bit for bit."""
import os

import numpy as np
import pytest

import _cases

pytestmark = pytest.mark.gpu
pkg = _cases.pkg
P = pkg.params
slab = pkg.slab

CASES = {
    "sc2d_contact": (lambda: P.sc_params(P.MODEL_SC_D2Q9, 48, 32, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT),
                     P.CASE_SC_CONTACT2D, (0.265, 0.038, 9.0), 61),
    "sc3d_sessile": (lambda: P.sc_params(P.MODEL_SC_D3Q19, 24, 16, 20, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT),
                     P.CASE_SC_DROPLET3D, (0.265, 0.038, 6.0, 5.0), 41),
    "hcz2d_rt": (lambda: P.hcz_params(P.MODEL_HCZ_D2Q9, 32, 66, N=32), P.CASE_HCZ_RT2D, (), 41),
    "hcz3d_drop": (lambda: P.hcz_params(P.MODEL_HCZ_D3Q19, 24, 12, 12, ulb=0.01, N=24, Re=6.0, kappa=5e-4, gravity=-1e-5),
                   P.CASE_HCZ_LAPLACE3D, (), 25),
}


def single_run(prm, case, args, steps):
    with pkg.clbm.Lattice(prm) as single:
        single.init_case(case, args)
        single.step(steps)
        return single.in_pops(), single.fields()


def make_ring(prm, case, args, nranks):
    lats = [pkg.clbm.Lattice(slab.slab_params(prm, r, nranks)) for r in range(nranks)]
    for lat in lats:
        lat.init_case(case, args)
    ring = slab.LocalRing(lats, peer=True)
    assert all(lat.ring_kind() == 2 for lat in lats)
    return lats, ring


@pytest.mark.parametrize("graph", [1, 0])
@pytest.mark.parametrize("nranks", [2, 3])
@pytest.mark.parametrize("name", sorted(CASES))
def test_peer_ring_matches_single_slab(name, nranks, graph):
    mk, case, args, steps = CASES[name]
    prm = mk()
    ref_pops, ref_fields = single_run(prm, case, args, steps)
    old = os.environ.get("CLBM_SLAB_GRAPH")
    os.environ["CLBM_SLAB_GRAPH"] = str(graph)          # read once per context, in clbm_create
    try:
        lats, ring = make_ring(prm, case, args, nranks)
    finally:
        if old is None:
            os.environ.pop("CLBM_SLAB_GRAPH", None)
        else:
            os.environ["CLBM_SLAB_GRAPH"] = old
    # odd step count, chunks of 7: eager first steps, graph replays (two steps each) and an eager odd step all occur
    ring.step(steps, chunk=7)
    ring.refresh_moment_halo()
    pops = np.concatenate([lat.in_pops() for lat in lats], axis=2)
    fields = {k: np.concatenate([lat.fields()[k] for lat in lats]) for k in ("s0", "ux", "uy", "uz")}
    for lat in lats:
        lat.close()
    # HCZ D3Q19: the single-sweep kernel sums the moments of an x-boundary plane in a different order than the slab's
    # boundary-plane pass; everything else is bit-identical
    if name == "hcz3d_drop":
        assert _cases.rel_linf(pops, ref_pops) < 1e-13
        for k in fields:
            assert _cases.rel_linf(fields[k], ref_fields[k]) < 1e-12, k
        return
    np.testing.assert_array_equal(pops, ref_pops)
    for k in fields:
        np.testing.assert_array_equal(fields[k], ref_fields[k])


def test_protocols_mix_without_host_synchronisation():
    """a step of the sequential protocol (stages 0-2 on the launching stream) between steps of the overlap protocol (stages
    10-12, boundary stream), every call asynchronous: the cross-stream ordering must come from events, not from a host sync
    (the boundary stream has to wait for what the launching stream ran last)"""
    prm = P.sc_params(P.MODEL_SC_D3Q19, 32, 24, 36, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT)
    case, args = P.CASE_SC_DROPLET3D, (0.265, 0.038, 8.0, 5.0)
    ref_pops, _ = single_run(prm, case, args, 30)
    lats, ring = make_ring(prm, case, args, 2)
    assert all(lat.overlap_supported() for lat in lats)

    def sequential_step():
        for st, ph in ((0, 0), (1, 1)):
            for lat in lats:
                lat.step_stage(st)
            for lat in lats:
                lat.slab_exchange(ph)
        for lat in lats:
            lat.step_stage(2)

    for _ in range(5):
        ring.step(3, chunk=3)          # overlap protocol (graph replays from the third step on)
        sequential_step()
        ring.step(1)
        sequential_step()
    pops = np.concatenate([lat.in_pops() for lat in lats], axis=2)
    for lat in lats:
        lat.close()
    np.testing.assert_array_equal(pops, ref_pops)


def test_peer_ring_full_plane_slabs_hcz2d_config3_shape():
    """BASELINE configs[2] plane shape: 4 slabs of 16 columns x 8194 rows, overlap protocol + graph replay, against the single
    64 x 8194 slab, bit for bit"""
    prm = P.hcz_params(P.MODEL_HCZ_D2Q9, 64, 8194, 1, ulb=0.04, N=2048, Re=3000.0)
    ref_pops, _ = single_run(prm, P.CASE_HCZ_RT2D, (), 12)
    lats, ring = make_ring(prm, P.CASE_HCZ_RT2D, (), 4)
    ring.step(12, chunk=4)
    pops = np.concatenate([lat.in_pops() for lat in lats], axis=2)
    for lat in lats:
        lat.close()
    np.testing.assert_array_equal(pops, ref_pops)
