"""GPU parity of the MRT collision operator of the HCZ D2Q9 kernels (CLBM_COLLISION_MRT) -- through the C ABI.

The operator has no reference implementation (the reference's HCZ functor is BGK; BASELINE's config text asks for MRT): parity
unpinned against the reference.  It is pinned (a) to the BGK kernels, which are parity-tested against the reference-pinned
oracle, at S = omega I, and (b) to the oracle's matrix-form MRT (CooLBM_MRT_combustion.cpp's M / M^-1 S M convention) at free
rates.  Bar: 1e-10 relative L-inf after up to 1000 steps; masks bit-exact."""
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


@pytest.mark.parametrize("fused", [0, 1])
def test_mrt_equal_rates_reproduces_the_bgk_kernels(fused):
    bgk = P.hcz_params(P.MODEL_HCZ_D2Q9, 64, 258, N=256)            # omega of BASELINE configs[1] (1.95986)
    mrt = P.hcz_mrt_params(64, 258, N=256)
    fa, pa, _ = _gpu(bgk, P.CASE_HCZ_RT2D, (), 300, fused)
    fb, pb, _ = _gpu(mrt, P.CASE_HCZ_RT2D, (), 300, fused)
    assert rel_linf(pb, pa) < 1e-12
    for k in ("s0", "s1", "s2", "ux", "uy"):
        assert rel_linf(fb[k], fa[k]) < 1e-11, k


_ORACLE_RUNS = {}


def _oracle_1000(rates):
    """one oracle run per parameter set, shared by the fused / staged parametrisations (box time)"""
    if rates not in _ORACLE_RUNS:
        prm = P.hcz_params(P.MODEL_HCZ_D2Q9, 48, 194, N=256) if rates is None else \
            P.hcz_mrt_params(48, 194, N=256, s_e=rates[0], s_eps=rates[1], s_q=rates[2])
        _ORACLE_RUNS[rates] = (prm, OracleSim(prm).init_case(P.CASE_HCZ_RT2D, ()).step(1000))
    return _ORACLE_RUNS[rates]


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("rates", [(1.8, 1.9, 1.7), (1.4, 1.6, 1.8)])   # stable at omega = 1.96 (several other triples blow up, in the oracle too)
def test_mrt_free_rates_match_the_oracle_1000_steps(rates, fused):
    prm, ora = _oracle_1000(rates)
    got, pops, flags = _gpu(prm, P.CASE_HCZ_RT2D, (), 1000, fused)
    np.testing.assert_array_equal(flags, ora.flag)
    ref = ora.fields()
    for k in ("s0", "s1", "s2", "ux", "uy"):
        assert rel_linf(got[k], ref[k]) < TOL, k
    assert rel_linf(pops, ora.in_pops()) < TOL
    bgk = _oracle_1000(None)[1].fields()
    assert rel_linf(ref["uy"], bgk["uy"]) > 1e-4          # the free rates do change the solution


@pytest.mark.parametrize("fused", [0, 1])
def test_mrt_layered_variant_matches_the_oracle(fused):
    prm = P.hcz_layered_params(10, 101, gx_const=1e-6).copy(collision=P.COLLISION_MRT, s_e=1.2, s_eps=1.1, s_q=1.4)
    args = (0.3, 2.0)
    got, pops, flags = _gpu(prm, P.CASE_HCZ_LAYERED2D, args, 500, fused)
    ora = OracleSim(prm).init_case(P.CASE_HCZ_LAYERED2D, args).step(500)
    ref = ora.fields()
    for k in ("s0", "s1", "s2", "ux", "uy"):
        if np.max(np.abs(ref[k])) > 0:
            assert rel_linf(got[k], ref[k]) < TOL, k
    assert rel_linf(pops, ora.in_pops()) < TOL


def test_mrt_odd_sizes_and_x_chunks():
    prm = P.hcz_mrt_params(100, 300, omega=1.8, s_e=1.3, s_eps=1.0, s_q=1.6)
    got, pops, _ = _gpu(prm, P.CASE_HCZ_RT2D, (), 150, 1)
    ora = OracleSim(prm).init_case(P.CASE_HCZ_RT2D, ()).step(150)
    assert rel_linf(pops, ora.in_pops()) < TOL


@pytest.mark.parametrize("fused", [0, 1])
def test_mrt_slab_ring_matches_single_slab(fused):
    slab = pkg.slab
    prm = P.hcz_mrt_params(32, 66, N=32, s_e=1.1, s_eps=1.2, s_q=1.3).copy(fused=fused)
    ora = OracleSim(prm).init_case(P.CASE_HCZ_RT2D, ())
    with pkg.clbm.Lattice(prm) as single:
        single.upload(ora.lattice, ora.flag, 0)
        single.step(80)
        ref_pops = single.in_pops()
    lats = []
    for r in range(2):
        lat = pkg.clbm.Lattice(slab.slab_params(prm, r, 2))
        l, f = slab.slice_host_state(prm, ora.lattice, ora.flag, r, 2)
        lat.upload(l, f, 0)
        lats.append(lat)
    ring = slab.LocalRing(lats)
    ring.exchange_flags()
    ring.step(80)
    pops = np.concatenate([lat.in_pops() for lat in lats], axis=2)
    for lat in lats:
        lat.close()
    np.testing.assert_array_equal(pops, ref_pops)


def test_mrt_rates_outside_the_stable_range_and_unknown_operators_are_rejected():
    for prm in (P.hcz_params(P.MODEL_HCZ_D3Q19, 8, 8, 8).copy(collision=2, s_e=1.0, s_eps=1.0, s_q=1.0),
                P.sc_params(P.MODEL_SC_D3Q19, 8, 8, 8).copy(collision=P.COLLISION_MRT, s_e=1.0, s_eps=1.0, s_q=2.0),
                P.hcz_mrt_params(16, 66, s_e=2.5), P.hcz_mrt_params(16, 66, s_q=0.0).copy(s_q=0.0)):
        with pytest.raises(pkg.clbm.ClbmError):
            pkg.clbm.Lattice(prm)


def test_mrt_speed_line(capsys):
    """not a parity check: prints the MLUPS of the MRT fused kernel next to BGK at 2048 x 8194 for DESIGN.md"""
    out = []
    for name, prm in (("bgk", P.hcz_params(P.MODEL_HCZ_D2Q9, 2048, 8194, N=2048)), ("mrt", P.hcz_mrt_params(2048, 8194, N=2048, s_e=1.6, s_eps=1.65, s_q=1.7))):
        with pkg.clbm.Lattice(prm) as lat:
            lat.init_case(P.CASE_HCZ_RT2D, ())
            lat.step(5)
            lat.sync()
            ms = lat.step_timed(40)
            assert np.isfinite(lat.reduce(P.REDUCE_MASS))
        out.append("%s %.0f MLUPS" % (name, prm.nelem * 40 / (ms * 1e3)))
    with capsys.disabled():
        print("\nHCZ D2Q9 2048x8194 fused: " + ", ".join(out))
