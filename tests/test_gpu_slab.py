"""x-slab decomposition on ONE GPU: R slab contexts in one process exchange their ghost planes with
device-to-device copies (LocalRing).  A slab run must reproduce the single-slab run bit-for-bit
(the arithmetic per node is identical; only the data path differs) and the oracle within 1e-10."""
import numpy as np
import pytest

import _cases
from _oracle import OracleSim

pytestmark = pytest.mark.gpu
pkg = _cases.pkg
P = pkg.params
slab = pkg.slab

CASES = {
    "sc2d_contact": (lambda: P.sc_params(P.MODEL_SC_D2Q9, 48, 32, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT),
                     P.CASE_SC_CONTACT2D, (0.265, 0.038, 9.0), 120),
    "sc3d_sessile": (lambda: P.sc_params(P.MODEL_SC_D3Q19, 24, 16, 20, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT),
                     P.CASE_SC_DROPLET3D, (0.265, 0.038, 6.0, 5.0), 60),
    "hcz2d_rt": (lambda: P.hcz_params(P.MODEL_HCZ_D2Q9, 32, 66, N=32), P.CASE_HCZ_RT2D, (), 80),
    "hcz3d_drop": (lambda: P.hcz_params(P.MODEL_HCZ_D3Q19, 24, 12, 12, ulb=0.01, N=24, Re=6.0, kappa=5e-4, gravity=-1e-5),
                   P.CASE_HCZ_LAPLACE3D, (), 40),
}


@pytest.mark.parametrize("fused", [0, 1, 2])
@pytest.mark.parametrize("nranks", [2, 3, 4])
@pytest.mark.parametrize("name", sorted(CASES))
def test_slab_ring_matches_single_slab(name, nranks, fused):
    mk, case, args, steps = CASES[name]
    prm = mk().copy(fused=fused)
    ora = OracleSim(prm).init_case(case, args)
    with pkg.clbm.Lattice(prm) as single:
        single.upload(ora.lattice, ora.flag, 0)
        single.step(steps)
        ref_pops = single.in_pops()
        ref_fields = single.fields()

    lats = []
    for r in range(nranks):
        lat = pkg.clbm.Lattice(slab.slab_params(prm, r, nranks))
        l, f = slab.slice_host_state(prm, ora.lattice, ora.flag, r, nranks)
        lat.upload(l, f, 0)
        lats.append(lat)
    ring = slab.LocalRing(lats)
    ring.exchange_flags()
    ring.step(steps)
    ring.refresh_moment_halo()
    pops = np.concatenate([lat.in_pops() for lat in lats], axis=2)
    fields = {k: np.concatenate([lat.fields()[k] for lat in lats]) for k in ("s0", "s1", "ux", "uy", "uz")}
    for lat in lats:
        lat.close()
    np.testing.assert_array_equal(pops, ref_pops)
    for k in fields:
        np.testing.assert_array_equal(fields[k], ref_fields[k])
    ora.step(steps)
    assert _cases.rel_linf(pops, ora.in_pops()) < 1e-10


def test_slab_device_init_has_consistent_ghost_flags():
    prm = P.sc_params(P.MODEL_SC_D3Q19, 16, 12, 8, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT)
    args = (0.265, 0.038, 4.0, 5.0)
    with pkg.clbm.Lattice(prm) as single:               # same kernel variant as the slabs: bit-identical
        single.init_case(P.CASE_SC_DROPLET3D, args)
        single.step(30)
        ref = single.in_pops()
    lats = [pkg.clbm.Lattice(slab.slab_params(prm, r, 2)) for r in range(2)]
    for lat in lats:
        lat.init_case(P.CASE_SC_DROPLET3D, args)     # ghost flags come from the global geometry, no exchange needed
    ring = slab.LocalRing(lats)
    ring.step(30)
    pops = np.concatenate([lat.in_pops() for lat in lats], axis=2)
    for lat in lats:
        lat.close()
    np.testing.assert_array_equal(pops, ref)


class overlap_env:
    """CLBM_SLAB_OVERLAP (read once per context in clbm_create): 0 sequential only, 1 interior-first, 2 halo-first overlap"""

    def __init__(self, form):
        self.form = form

    def __enter__(self):
        import os
        self.old = os.environ.get("CLBM_SLAB_OVERLAP")
        os.environ["CLBM_SLAB_OVERLAP"] = str(self.form)

    def __exit__(self, *a):
        import os
        if self.old is None:
            os.environ.pop("CLBM_SLAB_OVERLAP", None)
        else:
            os.environ["CLBM_SLAB_OVERLAP"] = self.old


@pytest.mark.parametrize("form", [1, 2])
@pytest.mark.parametrize("nranks", [2, 4])
def test_overlap_protocol_is_bit_identical(nranks, form):
    """both forms of the overlap protocol (clbm_step_stage 10-12; 1: interior planes on the launching stream while the boundary
    stream runs both exchanges and the boundary planes, 2: moment halo first, boundary chunks, interior overlapping the
    population exchange) against the sequential protocol (stages 0-2, the Shan-Chen default) and the single slab"""
    prm = P.sc_params(P.MODEL_SC_D3Q19, 32, 24, 36, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT)
    args = (0.265, 0.038, 8.0, 5.0)
    with pkg.clbm.Lattice(prm) as single:
        single.init_case(P.CASE_SC_DROPLET3D, args)
        single.step(50)
        ref = single.in_pops()
    out = {}
    for overlap in (False, True):
        with overlap_env(form):
            lats = [pkg.clbm.Lattice(slab.slab_params(prm, r, nranks)) for r in range(nranks)]
        assert all(lat.overlap_supported() and lat.overlap_variant() == form for lat in lats)
        for lat in lats:
            lat.init_case(P.CASE_SC_DROPLET3D, args)
        ring = slab.LocalRing(lats)
        ring.step(25, overlap=overlap)
        ring.step(25, overlap=not overlap)      # the two protocols can be mixed between steps
        out[overlap] = np.concatenate([lat.in_pops() for lat in lats], axis=2)
        for lat in lats:
            lat.close()
    np.testing.assert_array_equal(out[False], ref)
    np.testing.assert_array_equal(out[True], ref)


@pytest.mark.parametrize("form", [0, 1, 2])
def test_full_plane_slabs_match_single_slab(form):
    """four slabs of 8 planes at the production plane size 512 x 512 (real tile grid, TMA boxes, 24-plane chunk logic) against
    the single 32 x 512 x 512 slab, bit for bit, with the sequential protocol and both forms of the overlap protocol"""
    prm = P.sc_params(P.MODEL_SC_D3Q19, 32, 512, 512, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT)
    args = (0.265, 0.038, 12.0, 5.0)
    with pkg.clbm.Lattice(prm) as single:
        single.init_case(P.CASE_SC_DROPLET3D, args)
        single.step(6)
        ref = single.in_pops()
    with overlap_env(form):
        lats = [pkg.clbm.Lattice(slab.slab_params(prm, r, 4)) for r in range(4)]
    assert all(lat.overlap_variant() == form for lat in lats)
    for lat in lats:
        lat.init_case(P.CASE_SC_DROPLET3D, args)
    ring = slab.LocalRing(lats)
    ring.step(6)
    pops = np.concatenate([lat.in_pops() for lat in lats], axis=2)
    for lat in lats:
        lat.close()
    np.testing.assert_array_equal(pops, ref)
