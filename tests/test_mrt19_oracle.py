"""The MRT collision operators added for the D3Q19 models and the Shan-Chen Rayleigh-Taylor (Guo) variant, in the oracle.

The reference's SC / HCZ functors are BGK and its only MRT basis (CooLBM_MRT_combustion.cpp:313-347) is D2Q9, so these
operators have NO reference implementation: parity unpinned.  What pins them: the D3Q19 rows are the orthogonal basis of
d'Humieres et al. (2002) -- checked here: mutually orthogonal, first rows = the conserved moments -- and with S = omega I every
operator must reproduce the BGK oracle (pinned bit-for-bit to the untouched reference headers where one exists) to round-off."""
import ctypes

import numpy as np

import _cases
from _cases import P, rel_linf
from _oracle import OracleSim, lib


def _rows():
    M = np.zeros(361)
    n2 = np.zeros(19)
    dp = ctypes.POINTER(ctypes.c_double)
    lib().oracle_mrt19_rows(M.ctypes.data_as(dp), n2.ctypes.data_as(dp))
    return M.reshape(19, 19), n2


def test_d3q19_rows_are_orthogonal_and_start_with_the_conserved_moments():
    M, n2 = _rows()
    G = M @ M.T
    assert np.array_equal(G, np.diag(np.diag(G)))          # small integers: exact
    assert np.array_equal(np.diag(G), n2)
    assert np.array_equal(n2, [19, 2394, 252, 10, 40, 10, 40, 10, 40, 36, 72, 12, 24, 4, 4, 4, 8, 8, 8])
    c = np.array([[-1, 0, 0], [0, -1, 0], [0, 0, -1], [-1, -1, 0], [-1, 1, 0], [-1, 0, -1], [-1, 0, 1], [0, -1, -1], [0, -1, 1], [0, 0, 0],
                  [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, -1, 0], [1, 0, 1], [1, 0, -1], [0, 1, 1], [0, 1, -1]], dtype=float)
    assert np.array_equal(M[0], np.ones(19))
    for row, axis in ((3, 0), (5, 1), (7, 2)):
        assert np.array_equal(M[row], c[:, axis])
    # M^-1 = M^T diag(1 / norm2)
    assert np.allclose((M.T / n2) @ M, np.eye(19), atol=1e-15)


def _mrt(p, **rates):
    om = p.omega
    return p.copy(collision=P.COLLISION_MRT, s_e=rates.get("s_e", om), s_eps=rates.get("s_eps", om), s_q=rates.get("s_q", om))


def test_sc_d3q19_mrt_with_equal_rates_is_bgk():
    for prm, case, args in (
            (P.sc_params(P.MODEL_SC_D3Q19, 28, 20, 24, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT), P.CASE_SC_DROPLET3D, (0.265, 0.038, 7.0, 5.0)),
            (P.sc_params(P.MODEL_SC_D3Q19, 20, 22, 24, omega=1.3, gravity=-2e-5, sc_force=P.SC_FORCE_LAPLACE), P.CASE_SC_DROPLET3D_PER, (0.265, 0.038, 6.0))):
        a = OracleSim(prm).init_case(case, args).step(150)
        b = OracleSim(_mrt(prm)).init_case(case, args).step(150)
        bulk = a.flag == 1
        assert np.isfinite(a.in_pops()[..., bulk]).all()
        assert rel_linf(b.in_pops()[..., bulk], a.in_pops()[..., bulk]) < 1e-12


def test_hcz_d3q19_mrt_with_equal_rates_is_the_pinned_bgk_operator():
    prm = P.hcz_params(P.MODEL_HCZ_D3Q19, 16, 16, 16, ulb=0.01, N=16, Re=6.0, kappa=5e-4, gravity=-1e-5)
    a = OracleSim(prm).init_case(P.CASE_HCZ_LAPLACE3D, ()).step(120)
    b = OracleSim(_mrt(prm)).init_case(P.CASE_HCZ_LAPLACE3D, ()).step(120)
    assert rel_linf(b.in_pops(), a.in_pops()) < 1e-12


def test_sc_rt_mrt_with_equal_rates_is_the_pinned_bgk_operator():
    prm = P.sc_rt_params(32, 130, omega=1.3)
    # (this model amplifies one-ulp differences -- tests/test_gpu_zz_sc_rt2d.py measures it -- hence 100 steps and 1e-11)
    a = OracleSim(prm).init_case(P.CASE_SC_RT2D, (1.2, 0.4)).step(100)
    b = OracleSim(_mrt(prm)).init_case(P.CASE_SC_RT2D, (1.2, 0.4)).step(100)
    bulk = a.flag == 1
    assert rel_linf(b.in_pops()[..., bulk], a.in_pops()[..., bulk]) < 1e-11


def test_free_rates_conserve_mass_and_are_not_a_no_op():
    prm = P.sc_params(P.MODEL_SC_D3Q19, 20, 22, 24, omega=1.3, gravity=-2e-5, sc_force=P.SC_FORCE_LAPLACE)
    a = OracleSim(prm).init_case(P.CASE_SC_DROPLET3D_PER, (0.265, 0.038, 6.0))
    c = OracleSim(_mrt(prm, s_e=1.1, s_eps=1.2, s_q=1.5)).init_case(P.CASE_SC_DROPLET3D_PER, (0.265, 0.038, 6.0))
    bulk = a.flag == 1
    m0 = c.in_pops()[0][:, bulk].sum()
    a.step(150)
    c.step(150)
    assert abs(c.in_pops()[0][:, bulk].sum() - m0) <= 1e-12 * abs(m0)
    d = rel_linf(c.fields()["s0"], a.fields()["s0"])
    assert np.isfinite(d) and 1e-7 < d < 0.2
