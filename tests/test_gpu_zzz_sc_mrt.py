"""GPU parity of the MRT collision operator for Yuan-CS Shan-Chen D2Q9 (CLBM_COLLISION_MRT) -- through the C ABI.

No reference implementation exists (the reference's Shan-Chen functors are BGK): parity unpinned against the reference; pinned
to the BGK kernels at S = omega I and to the oracle's matrix-form MRT at free rates (1e-10).  The per-cell arithmetic
(sc_collide_mrt) is also checked on the CPU (tests/test_host_check.py).  NOTE: written after the round's GPU budget was spent --
these tests have not run on a GPU yet, which is why the file sorts last."""
import numpy as np
import pytest

import _cases
from _cases import rel_linf
from _oracle import OracleSim

pytestmark = pytest.mark.gpu

pkg = _cases.pkg
P = pkg.params
TOL = 1e-10


def _gpu(prm, case, args, steps, fused):
    with pkg.clbm.Lattice(prm.copy(fused=fused)) as lat:
        lat.init_case(case, args)
        lat.step(steps)
        return lat.fields(), lat.in_pops(), lat.flags()


CASES = {
    "laplace": (lambda **k: P.sc_params(P.MODEL_SC_D2Q9, 96, 80, omega=1.2, gravity=-1e-5, **k), P.CASE_SC_LAPLACE2D, (0.265, 0.038, 10.0)),
    "contact": (lambda **k: P.sc_params(P.MODEL_SC_D2Q9, 96, 48, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT, **k), P.CASE_SC_CONTACT2D, (0.265, 0.038, 14.0)),
    "layered": (lambda **k: P.sc_layered_params(10, 101, omega=1.1, gx=1e-6, **k), P.CASE_SC_LAYERED2D, (0.21, 0.067, 0.3, 4.0)),
}


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("name", sorted(CASES))
def test_sc_mrt_equal_rates_reproduces_the_bgk_kernels(name, fused):
    mk, case, args = CASES[name]
    bgk = mk()
    mrt = bgk.copy(collision=P.COLLISION_MRT, s_e=bgk.omega, s_eps=bgk.omega, s_q=bgk.omega)
    _, pa, _ = _gpu(bgk, case, args, 300, fused)
    _, pb, _ = _gpu(mrt, case, args, 300, fused)
    assert rel_linf(pb, pa) < 1e-11


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("name", sorted(CASES))
def test_sc_mrt_free_rates_match_the_oracle_1000_steps(name, fused):
    mk, case, args = CASES[name]
    bgk = mk()
    prm = bgk.copy(collision=P.COLLISION_MRT, s_e=min(1.9, bgk.omega + 0.3), s_eps=max(0.5, bgk.omega - 0.2), s_q=1.4)
    got, pops, flags = _gpu(prm, case, args, 1000, fused)
    ora = OracleSim(prm).init_case(case, args).step(1000)
    np.testing.assert_array_equal(flags, ora.flag)
    ref = ora.fields()
    for k in ("s0", "s1", "ux", "uy"):
        if np.max(np.abs(ref[k])) > 0:
            assert rel_linf(got[k], ref[k]) < TOL, k
    assert rel_linf(pops, ora.in_pops()) < TOL


def test_sc_mrt_slab_ring_matches_single_slab():
    slab = pkg.slab
    mk, case, args = CASES["contact"]
    prm = mk().copy(collision=P.COLLISION_MRT, s_e=1.3, s_eps=0.8, s_q=1.4)
    ora = OracleSim(prm).init_case(case, args)
    with pkg.clbm.Lattice(prm) as single:
        single.upload(ora.lattice, ora.flag, 0)
        single.step(80)
        ref_pops = single.in_pops()
    lats = []
    for r in range(3):
        lat = pkg.clbm.Lattice(slab.slab_params(prm, r, 3))
        l, f = slab.slice_host_state(prm, ora.lattice, ora.flag, r, 3)
        lat.upload(l, f, 0)
        lats.append(lat)
    ring = slab.LocalRing(lats)
    ring.exchange_flags()
    ring.step(80)
    pops = np.concatenate([lat.in_pops() for lat in lats], axis=2)
    for lat in lats:
        lat.close()
    np.testing.assert_array_equal(pops, ref_pops)


def test_sc_mrt_exists_for_every_shan_chen_variant_now():
    """round 2 added the D3Q19 operator and the Rayleigh-Taylor / Guo one (tests/test_gpu_zzzz_mrt19.py)"""
    for prm in (P.sc_params(P.MODEL_SC_D3Q19, 8, 8, 8).copy(collision=P.COLLISION_MRT, s_e=1.0, s_eps=1.0, s_q=1.0),
                P.sc_rt_params(16, 66).copy(collision=P.COLLISION_MRT, s_e=1.0, s_eps=1.0, s_q=1.0)):
        with pkg.clbm.Lattice(prm) as lat:
            assert lat is not None
