"""GPU parity of the Shan-Chen Rayleigh-Taylor variant (SC/apps/RayleighTaylor2D.h: psi = 1 - exp(-rho), mirrored psi at
wall neighbours, Guo forcing; CLBM_SC_FORCE_EXPGUO) -- through the C ABI, the CPU oracle is only the checker.

Bar: 1e-10 relative L-inf.  Scalars and populations use the global-max norm; velocity and force are normalised as VECTOR
fields (largest component, _cases.rel_linf_vec): in this case the x components decay to ~1e-6 while y stays O(0.1), so a
per-component norm would measure round-off against round-off.  Masks bit-exact.
"""
import numpy as np
import pytest

import _cases
from _cases import rel_linf, rel_linf_vec
from _oracle import OracleSim

pytestmark = pytest.mark.gpu

pkg = _cases.pkg
P = pkg.params
TOL = 1e-10


def _run(prm, args, steps, fused, device_init=False):
    ora = OracleSim(prm).init_case(P.CASE_SC_RT2D, args)
    with pkg.clbm.Lattice(prm.copy(fused=fused)) as lat:
        if device_init:
            lat.init_case(P.CASE_SC_RT2D, args)
        else:
            lat.upload(ora.lattice, ora.flag, 0)
        lat.step(steps)
        got, pops, flags, F = lat.fields(), lat.in_pops(), lat.flags(), lat.force()
        energy, mass = lat.reduce(P.REDUCE_ENERGY), lat.reduce(P.REDUCE_MASS)
    ora.step(steps)
    return ora, got, pops, flags, F, energy, mass


def _check(ora, got, pops, flags, F, ref=None, refF=None):
    golden = ref is not None
    ref = ref or ora.fields()
    refF = refF or ora.force()
    np.testing.assert_array_equal(flags, ora.flag)
    assert rel_linf(got["s0"], ref["s0"]) < TOL
    # P_eos (:200-208, a diagnostic the reference defines but never writes) has a POLE at rho = 4/b, inside [rhog, rhol] of the
    # shipped parameters (b = 4: rho = 1), so it amplifies the (1e-11-level, after 1000 steps of this violent start-up flow)
    # density differences without bound.  It is therefore checked as a FORMULA: the device value against the same expression
    # evaluated here from the device's own density (which is itself held to 1e-10 above), and against the reference's dump
    # where that was produced by the untouched header (golden fixtures, away from the pole).
    r = got["s0"]
    rt = ora.p.b * r / 4.0
    with np.errstate(divide="ignore", invalid="ignore"):
        t1, t2 = (r / 3.0) * (1.0 + rt + rt * rt - rt * rt * rt) / ((1 - rt) * (1 - rt) * (1 - rt)), ora.p.a * r * r
    bulk = ora.flag == 1
    assert np.all(np.abs(got["s1"][bulk] - (t1 - t2)[bulk]) <= 1e-12 * (np.abs(t1) + np.abs(t2))[bulk])
    assert np.all(got["s1"][~bulk] == 0.0)
    if golden:
        far = np.abs(1.0 - ora.p.b * ref["s0"] / 4.0) > 0.25
        assert rel_linf(got["s1"][far], ref["s1"][far]) < TOL
    assert rel_linf_vec([got["ux"], got["uy"]], [ref["ux"], ref["uy"]]) < TOL
    assert rel_linf_vec([F[0], F[1]], [refF["fx"], refF["fy"]]) < TOL
    assert np.max(np.abs(got["uz"])) == 0.0
    assert rel_linf(pops, ora.in_pops()) < TOL


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("name", _cases.golden_names("sc_rt2d"))
def test_sc_rt2d_golden_fixture(name, fused):
    """fixtures dumped by the untouched reference header (populations, density, P_eos, u_eq, force_ff)"""
    z, _ = _cases.load_golden(name)
    prm, case, args, steps, _ = _cases.golden_setup(name)
    ora, got, pops, flags, F, _, _ = _run(prm, args, steps, fused)
    np.testing.assert_array_equal(flags, z["flag"])
    _check(ora, got, pops, flags, F, ref={"s0": z["rho"], "s1": z["pressure"], "ux": z["ux"], "uy": z["uy"]},
           refF={"fx": z["fx"], "fy": z["fy"]})
    assert rel_linf(pops, z["pops"]) < TOL


def _ulp_sensitivity(prm, args, steps):
    """response of the ORACLE itself to a one-ulp relative perturbation of its initial populations, after `steps` steps:
    how much of an error any implementation that rounds differently (FMA contraction, another exp()) must be expected to show"""
    a = OracleSim(prm).init_case(P.CASE_SC_RT2D, args)
    b = OracleSim(prm).init_case(P.CASE_SC_RT2D, args)
    rng = np.random.default_rng(0)
    b.lattice *= 1.0 + rng.integers(-1, 2, size=b.lattice.size) * 2.2e-16
    a.step(steps)
    b.step(steps)
    fa, fb = a.fields(), b.fields()
    return max(rel_linf(fb["s0"], fa["s0"]), rel_linf_vec([fb["ux"], fb["uy"]], [fa["ux"], fa["uy"]]), rel_linf(b.in_pops(), a.in_pops()))


@pytest.mark.parametrize("fused", [0, 1])
def test_sc_rt2d_200_steps(fused):
    """shipped parameters (config_RayleighTaylor2D.txt: omega = 1, g = -5, gravity = -1.25e-5) on 64 x 258, device-side
    initial condition, 200 steps (past the start-up transient, before the amplification documented below), 1e-10 on every field"""
    prm = P.sc_rt_params(64, omega=1.0)
    ora, got, pops, flags, F, energy, mass = _run(prm, (1.2, 0.4), 200, fused, device_init=True)
    _check(ora, got, pops, flags, F)
    ref = ora.fields()
    bulk = ora.flag == 1
    e_ref = 0.5 * np.sum((ref["ux"] ** 2 + ref["uy"] ** 2)[bulk]) / prm.nelem
    assert abs(energy - e_ref) <= 1e-10 * e_ref
    assert abs(mass - ref["s0"][bulk].sum()) <= 1e-12 * mass


@pytest.mark.parametrize("fused", [0, 1])
def test_sc_rt2d_1000_steps(fused):
    """the north_star horizon.  This case starts from a tanh profile far from the psi = 1 - exp(-rho) equilibrium (velocities
    up to 0.3 lattice units) and AMPLIFIES rounding: the untouched model answers a one-ulp perturbation of its initial
    populations with 2e-10 (velocity) after 1000 steps and 1e-9 after 500 (measured with the oracle, _ulp_sensitivity).  The
    bar here is therefore the larger of 1e-10 and 10x that self-sensitivity; measured on the B200: 1.8e-10."""
    prm = P.sc_rt_params(64, omega=1.0)
    args = (1.2, 0.4)
    bar = max(TOL, 10.0 * _ulp_sensitivity(prm, args, 1000))
    assert bar < 1e-8
    ora, got, pops, flags, F, _, _ = _run(prm, args, 1000, fused, device_init=True)
    ref, refF = ora.fields(), ora.force()
    np.testing.assert_array_equal(flags, ora.flag)
    assert rel_linf(got["s0"], ref["s0"]) < bar
    assert rel_linf_vec([got["ux"], got["uy"]], [ref["ux"], ref["uy"]]) < bar
    assert rel_linf_vec([F[0], F[1]], [refF["fx"], refF["fy"]]) < bar
    assert rel_linf(pops, ora.in_pops()) < bar


def test_sc_rt2d_odd_sizes_and_chunks():
    """partial tiles of the fused kernel (ny not a multiple of 128) and more than one x-chunk"""
    prm = P.sc_rt_params(70, 150, omega=1.4, g=-4.5, gravity=-5e-5)
    ora, got, pops, flags, F, _, _ = _run(prm, (1.5, 0.3), 200, 1)
    _check(ora, got, pops, flags, F)


def test_sc_rt2d_mass_conservation_large():
    """size-independent property at a lattice the oracle does not run: bounce-back walls + periodic x conserve mass"""
    prm = P.sc_rt_params(512, omega=1.0)
    with pkg.clbm.Lattice(prm) as lat:
        lat.init_case(P.CASE_SC_RT2D, (1.2, 0.4))
        m0 = lat.reduce(P.REDUCE_MASS)
        lat.step(500)
        m1 = lat.reduce(P.REDUCE_MASS)
    assert abs(m1 - m0) <= 1e-11 * m0


def test_sc_rt2d_rejected_on_d3q19():
    prm = P.sc_params(P.MODEL_SC_D3Q19, 8, 8, 8, sc_force=P.SC_FORCE_EXPGUO)
    with pytest.raises(pkg.clbm.ClbmError):
        pkg.clbm.Lattice(prm)


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("nranks", [2, 3])
def test_sc_rt2d_slab_ring_matches_single_slab(nranks, fused):
    """x-slab decomposition (psi halo of depth 1, crossing populations): bit-identical to the single slab"""
    slab = pkg.slab
    prm = P.sc_rt_params(24, 50, omega=1.2).copy(fused=fused)
    args, steps = (1.2, 0.4), 80
    ora = OracleSim(prm).init_case(P.CASE_SC_RT2D, args)
    with pkg.clbm.Lattice(prm) as single:
        single.upload(ora.lattice, ora.flag, 0)
        single.step(steps)
        ref_pops = single.in_pops()
    lats = []
    for r in range(nranks):
        lat = pkg.clbm.Lattice(slab.slab_params(prm, r, nranks))
        l, f = slab.slice_host_state(prm, ora.lattice, ora.flag, r, nranks)
        lat.upload(l, f, 0)
        lats.append(lat)
    ring = slab.LocalRing(lats)
    ring.exchange_flags()
    ring.step(steps)
    pops = np.concatenate([lat.in_pops() for lat in lats], axis=2)
    for lat in lats:
        lat.close()
    np.testing.assert_array_equal(pops, ref_pops)
