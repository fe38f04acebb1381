"""L2-resident Shan-Chen D2Q9 lattices (BASELINE configs[0]): clbm_step(n) runs all n steps in one cooperative launch (sc_fused.cu:
the column-resident kernel with a grid barrier, or the plane-marching kernel in its MULTI form when the columns do not fit).
Both must give populations bit-identical to the launch-by-launch path, for odd and even n, walls, gravity, several calls in a
row; and the oracle comparison of configs[0] itself at 1000 steps (1e-10) goes through the new path."""
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
    "c1_laplace_256": (P.sc_params(P.MODEL_SC_D2Q9, 256, 256, ulb=0.01, N=256, Re=6.0), P.CASE_SC_LAPLACE2D, (0.265, 0.038, 10.0)),
    "gravity_200x130": (P.sc_params(P.MODEL_SC_D2Q9, 200, 130, omega=1.2, gravity=-1e-5), P.CASE_SC_LAPLACE2D, (0.265, 0.038, 20.0)),
    "contact_walls_96x48": (P.sc_params(P.MODEL_SC_D2Q9, 96, 48, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT), P.CASE_SC_CONTACT2D, (0.265, 0.038, 14.0)),
    "contact_walls_512x256": (P.sc_params(P.MODEL_SC_D2Q9, 512, 256, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT), P.CASE_SC_CONTACT2D, (0.265, 0.038, 40.0)),
    "layered_constg_10x101": (P.sc_layered_params(10, 101, omega=1.1, gx=1e-6), P.CASE_SC_LAYERED2D, (0.21, 0.067, 0.3, 4.0)),
}


def _run(prm, case, args, calls, multi):
    if multi is None:
        os.environ.pop("CLBM_SC_MULTI", None)
    else:
        os.environ["CLBM_SC_MULTI"] = str(multi)
    try:
        with pkg.clbm.Lattice(prm) as lat:
            lat.init_case(case, args)
            l0 = lat.launch_count()
            for n in calls:
                lat.step(n)
            return lat.in_pops(), lat.launch_count() - l0
    finally:
        os.environ.pop("CLBM_SC_MULTI", None)


@pytest.mark.parametrize("name", sorted(CASES))
def test_multi_step_launch_is_bit_identical_to_single_steps(name):
    prm, case, args = CASES[name]
    calls = (1, 7, 40, 3, 2, 50)
    ref, n_ref = _run(prm, case, args, calls, 0)
    assert n_ref >= sum(calls)
    got, n_got = _run(prm, case, args, calls, None)
    np.testing.assert_array_equal(got, ref)
    assert n_got < n_ref            # the default really took the multi-step path
    for forced in (2, 6):           # plane-marching MULTI form; column-resident kernel with neighbour flags
        got, _ = _run(prm, case, args, calls, forced)
        np.testing.assert_array_equal(got, ref)


def test_config1_1000_steps_against_the_oracle_through_the_multi_step_path():
    prm, case, args = CASES["c1_laplace_256"]
    with pkg.clbm.Lattice(prm) as lat:
        lat.init_case(case, args)
        l0 = lat.launch_count()
        lat.step(1000)
        assert lat.launch_count() - l0 == 1
        got, pops = lat.fields(), lat.in_pops()
    ora = OracleSim(prm).init_case(case, args).step(1000)
    ref = ora.fields()
    for k in ("s0", "s1"):
        assert rel_linf(got[k], ref[k]) < 1e-10, k
    assert rel_linf(np.stack([got["ux"], got["uy"]]), np.stack([ref["ux"], ref["uy"]])) < 1e-10
    assert rel_linf(pops, ora.in_pops()) < 1e-10
