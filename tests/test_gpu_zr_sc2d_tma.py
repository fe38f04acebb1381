"""The TMA-staged D2Q9 Shan-Chen kernel (sc2d_tma.cu) against the register-pipelined one (sc_fused.cu): same per-cell functions and
summation order, so the populations must be BIT-IDENTICAL -- for every tile / stage shape, ragged last tiles, walls with contact
angle, gravity, the constant-G variant, short x-chunks -- and therefore inherit its oracle parity (test_gpu_parity.py); one direct
oracle comparison at 1000 steps on top."""
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
    "laplace_gravity_96x320": (P.sc_params(P.MODEL_SC_D2Q9, 96, 320, omega=1.2, gravity=-1e-5), P.CASE_SC_LAPLACE2D, (0.265, 0.038, 20.0)),
    "contact_walls_80x300": (P.sc_params(P.MODEL_SC_D2Q9, 80, 300, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT), P.CASE_SC_CONTACT2D, (0.265, 0.038, 30.0)),
    "laplace_64x64": (P.sc_params(P.MODEL_SC_D2Q9, 64, 64, ulb=0.01, N=64, Re=6.0), P.CASE_SC_LAPLACE2D, (0.265, 0.038, 10.0)),
    "rayleigh_taylor_guo_40x162": (P.sc_rt_params(40, 162, omega=1.3), P.CASE_SC_RT2D, (1.2, 0.4)),
    "layered_constg_10x102": (P.sc_layered_params(10, 102, omega=1.1, gx=1e-6), P.CASE_SC_LAYERED2D, (0.21, 0.067, 0.3, 4.0)),
}


def _run(prm, case, args, steps, tma, xchunk=None):
    os.environ["CLBM_SC2D_TMA"] = str(tma)
    os.environ["CLBM_SC_MULTI"] = "0"          # launch by launch: the kernel under test is the single-step one
    if xchunk:
        os.environ["CLBM_SC_XCHUNK"] = str(xchunk)
    try:
        with pkg.clbm.Lattice(prm) as lat:
            lat.init_case(case, args)
            lat.step(steps)
            return lat.in_pops(), lat.fields()
    finally:
        for k in ("CLBM_SC2D_TMA", "CLBM_SC_MULTI", "CLBM_SC_XCHUNK"):
            os.environ.pop(k, None)


@pytest.mark.parametrize("name", sorted(CASES))
def test_tma_kernel_is_bit_identical_to_the_register_pipelined_kernel(name):
    prm, case, args = CASES[name]
    ref, _ = _run(prm, case, args, 150, 0)
    for shape in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15):
        got, _ = _run(prm, case, args, 150, shape)
        np.testing.assert_array_equal(got, ref, err_msg="shape %d" % shape)
    for xchunk in (1, 3, 7):
        got, _ = _run(prm, case, args, 150, 1, xchunk)
        np.testing.assert_array_equal(got, ref, err_msg="x-chunk %d" % xchunk)


def test_tma_kernel_against_the_oracle_1000_steps():
    prm, case, args = CASES["contact_walls_80x300"]
    pops, got = _run(prm, case, args, 1000, 1)
    ora = OracleSim(prm).init_case(case, args).step(1000)
    ref = ora.fields()
    for k in ("s0", "s1"):
        assert rel_linf(got[k], ref[k]) < 1e-10, k
    assert rel_linf(np.stack([got["ux"], got["uy"]]), np.stack([ref["ux"], ref["uy"]])) < 1e-10
    bulk = ora.flag == 1
    assert rel_linf(pops[..., bulk], ora.in_pops()[..., bulk]) < 1e-10
