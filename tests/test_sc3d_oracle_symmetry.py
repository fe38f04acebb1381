"""Shan-Chen D3Q19 has no reference functor: its oracle is a composition (D3Q19 set of PF/apps/laplace3D.h + the
force / BGK of SC/apps/contactAngle2D.h), pinned indirectly.  Besides the z-uniform == D2Q9 test (GPU suite), this checks a
property a correct D3Q19 composition must have: the direction set, weights, force and streaming are invariant under swapping
the x and z axes, so a run on the transposed lattice is the transpose of the run (to round-off: summation orders differ)."""
import numpy as np
import pytest

import _cases
from _cases import P, rel_linf
from _oracle import OracleSim


@pytest.mark.parametrize("case,force", [(P.CASE_SC_DROPLET3D_PER, P.SC_FORCE_LAPLACE), (P.CASE_SC_DROPLET3D, P.SC_FORCE_CONTACT)])
def test_sc_d3q19_oracle_commutes_with_an_x_z_transpose(case, force):
    nx, ny, nz = 12, 10, 16
    args = (0.265, 0.038, 3.5, 4.0)
    a = OracleSim(P.sc_params(P.MODEL_SC_D3Q19, nx, ny, nz, tau=1.0, rho_w=0.2, sc_force=force)).init_case(case, args)
    b = OracleSim(P.sc_params(P.MODEL_SC_D3Q19, nz, ny, nx, tau=1.0, rho_w=0.2, sc_force=force)).init_case(case, args)
    # the droplet sits at (nx/2, yc, nz/2): the initial density of b is the transpose of a's
    ra = a.fields()["s0"].reshape(nx, ny, nz)
    rb = b.fields()["s0"].reshape(nz, ny, nx)
    np.testing.assert_array_equal(rb, ra.transpose(2, 1, 0))
    a.step(40)
    b.step(40)
    fa, fb = a.fields(), b.fields()
    T = lambda v, shape: v.reshape(shape)
    assert rel_linf(T(fb["s0"], (nz, ny, nx)), T(fa["s0"], (nx, ny, nz)).transpose(2, 1, 0)) < 1e-13
    assert rel_linf(T(fb["s1"], (nz, ny, nx)), T(fa["s1"], (nx, ny, nz)).transpose(2, 1, 0)) < 1e-12
    assert rel_linf(T(fb["uy"], (nz, ny, nx)), T(fa["uy"], (nx, ny, nz)).transpose(2, 1, 0)) < 1e-10
    # the x and z velocity components swap roles
    assert rel_linf(T(fb["ux"], (nz, ny, nx)), T(fa["uz"], (nx, ny, nz)).transpose(2, 1, 0)) < 1e-10
    assert rel_linf(T(fb["uz"], (nz, ny, nx)), T(fa["ux"], (nx, ny, nz)).transpose(2, 1, 0)) < 1e-10
    assert np.max(np.abs(fa["ux"])) > 1e-8
