"""The MRT collision operator of the HCZ D2Q9 model (include/clbm.h, CLBM_COLLISION_MRT) in the oracle.

BASELINE.json's configs name an MRT HCZ model; the reference's HCZ functor is BGK only, so this operator has
NO reference implementation: parity unpinned.  What pins it: evaluated with S = omega I it must reproduce the BGK oracle (which
is pinned bit-for-bit to the untouched reference header) to round-off, for the Rayleigh-Taylor and the layered variant."""
import numpy as np

import _cases
from _cases import P, rel_linf
from _oracle import OracleSim


def test_mrt_with_equal_rates_is_the_pinned_bgk_operator():
    for om in (1.0, 1.9):
        a = OracleSim(P.hcz_params(P.MODEL_HCZ_D2Q9, 24, 98, omega=om)).init_case(P.CASE_HCZ_RT2D, ()).step(400)
        b = OracleSim(P.hcz_mrt_params(24, 98, omega=om)).init_case(P.CASE_HCZ_RT2D, ()).step(400)
        assert rel_linf(b.in_pops(), a.in_pops()) < 1e-13


def test_mrt_layered_variant_with_equal_rates_is_bgk():
    pa = P.hcz_layered_params(10, 41, omega=1.2, gx_const=1e-6)
    pm = pa.copy(collision=P.COLLISION_MRT, s_e=1.2, s_eps=1.2, s_q=1.2)
    a = OracleSim(pa).init_case(P.CASE_HCZ_LAYERED2D, (0.3, 2.0)).step(200)
    b = OracleSim(pm).init_case(P.CASE_HCZ_LAYERED2D, (0.3, 2.0)).step(200)
    bulk = a.flag == 1
    assert rel_linf(b.in_pops()[..., bulk], a.in_pops()[..., bulk]) < 1e-13


def test_mrt_free_rates_conserve_mass_and_change_only_the_ghost_modes():
    """rates of e, eps, q differ from omega: the order parameter sum_k f_k is conserved to round-off (first row of M; the zeroth
    moment of g is not a conserved quantity of the HCZ model even with BGK), the solution stays finite and departs from BGK
    (the operator is not a no-op)"""
    om = 1.7
    a = OracleSim(P.hcz_params(P.MODEL_HCZ_D2Q9, 24, 98, omega=om)).init_case(P.CASE_HCZ_RT2D, ())
    c = OracleSim(P.hcz_mrt_params(24, 98, omega=om, s_e=1.1, s_eps=1.2, s_q=1.3)).init_case(P.CASE_HCZ_RT2D, ())
    bulk = a.flag == 1
    m0 = c.in_pops()[0][:, bulk].sum()
    a.step(300)
    c.step(300)
    m1 = c.in_pops()[0][:, bulk].sum()
    assert abs(m1 - m0) <= 1e-12 * abs(m0)
    fa, fc = a.fields(), c.fields()
    assert np.isfinite(fc["uy"]).all()
    assert 1e-4 < rel_linf(fc["s0"], fa["s0"]) < 0.2
