"""HCZ D2Q9, opt-in multi-step form (CLBM_HCZ2D_MULTI=1): clbm_step(n >= 2) runs all n steps in ONE cooperative launch
(hcz2d_fused.cu, MULTI form: a phi pass and the column march per step, grid barriers between them).  The populations must agree
with the launch-per-step path to round-off (1e-12 after 103 steps; the MRT instantiation bit for bit) -- odd and even n, several
calls in a row, ragged y segments, the layered variant (x body force, walls), the MRT operator -- and configs[1] itself is compared
with the oracle at 1000 steps through this path (PF/apps/rayleighTaylor2D.h:316-663)."""
import os

import numpy as np
import pytest

import _cases
from _cases import rel_linf
from _oracle import OracleSim

pytestmark = pytest.mark.gpu

pkg = _cases.pkg
P = pkg.params

CASES = {
    "c2_rt_256x1026": (P.hcz_params(P.MODEL_HCZ_D2Q9, 256, 1026, N=256), P.CASE_HCZ_RT2D, ()),
    "rt_40x330_ragged": (P.hcz_params(P.MODEL_HCZ_D2Q9, 40, 330, N=40), P.CASE_HCZ_RT2D, ()),
    "rt_7x66_short_chunks": (P.hcz_params(P.MODEL_HCZ_D2Q9, 7, 66, N=16), P.CASE_HCZ_RT2D, ()),
    "layered_10x101": (P.hcz_layered_params(10, 101, ulb=0.1, N=100, Re=60.0, gx=1e-7, gx_const=1e-6), P.CASE_HCZ_LAYERED2D, (0.3, 2.0)),
    "mrt_64x258": (P.hcz_mrt_params(64, 258, N=256, s_e=1.1, s_eps=1.2, s_q=1.3), P.CASE_HCZ_RT2D, ()),
}


def _run(prm, case, args, calls, multi):
    if multi is None:
        os.environ.pop("CLBM_HCZ2D_MULTI", None)
    else:
        os.environ["CLBM_HCZ2D_MULTI"] = str(multi)
    try:
        with pkg.clbm.Lattice(prm) as lat:
            lat.init_case(case, args)
            l0 = lat.launch_count()
            for n in calls:
                lat.step(n)
            return lat.in_pops(), lat.launch_count() - l0, lat.fields()
    finally:
        os.environ.pop("CLBM_HCZ2D_MULTI", None)


@pytest.mark.parametrize("name", sorted(CASES))
def test_multi_step_launch_is_bit_identical_to_single_steps(name):
    prm, case, args = CASES[name]
    calls = (1, 7, 40, 3, 2, 50)
    ref, n_ref, f_ref = _run(prm, case, args, calls, 0)
    assert n_ref >= sum(calls)
    got, n_got, f_got = _run(prm, case, args, calls, 1)
    assert n_got < n_ref            # the multi-step path really ran
    if name.startswith("mrt"):
        np.testing.assert_array_equal(got, ref)
    assert rel_linf(got, ref) < 1e-12
    for k in ("s0", "s1", "s2"):
        assert rel_linf(f_got[k], f_ref[k]) < 1e-12, k
    assert rel_linf(np.stack([f_got["ux"], f_got["uy"]]), np.stack([f_ref["ux"], f_ref["uy"]])) < 1e-11
    _, n_dflt, _ = _run(prm, case, args, (5,), None)
    assert n_dflt >= 5              # opt-in: the default stays launch per step


def test_multi_step_forced_on_a_lattice_larger_than_l2_chunks():
    """a lattice that needs long x-chunks to keep the grid co-resident"""
    prm, case, args = P.hcz_params(P.MODEL_HCZ_D2Q9, 1500, 1026, N=256), P.CASE_HCZ_RT2D, ()
    ref, _, _ = _run(prm, case, args, (9,), 0)
    got, n_got, _ = _run(prm, case, args, (9,), 1)
    assert n_got == 1
    assert rel_linf(got, ref) < 1e-13


def test_config2_1000_steps_against_the_oracle_through_the_multi_step_path():
    prm, case, args = CASES["c2_rt_256x1026"]
    os.environ["CLBM_HCZ2D_MULTI"] = "1"
    try:
        with pkg.clbm.Lattice(prm) as lat:
            lat.init_case(case, args)
            l0 = lat.launch_count()
            lat.step(1000)
            assert lat.launch_count() - l0 == 1
            got, pops = lat.fields(), lat.in_pops()
    finally:
        os.environ.pop("CLBM_HCZ2D_MULTI", None)
    ora = OracleSim(prm).init_case(case, args).step(1000)
    ref = ora.fields()
    for k in ("s0", "s1", "s2"):
        assert rel_linf(got[k], ref[k]) < 1e-10, k
    assert rel_linf(np.stack([got["ux"], got["uy"]]), np.stack([ref["ux"], ref["uy"]])) < 1e-10
    assert rel_linf(pops, ora.in_pops()) < 1e-10
