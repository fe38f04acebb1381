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


@pytest.mark.parametrize("axis", ["z", "x"])
def test_sc_d3q19_oracle_projects_onto_the_d2q9_reference_model(axis):
    """the composed D3Q19 oracle on a state that is uniform along z (resp. x) must equal the D2Q9 oracle -- which IS pinned
    bit-for-bit to the reference's contactAngle2D functor -- to round-off: the D3Q19 weights project onto the D2Q9 ones
    (1/18 + 2/36 = 1/9, 1/3 + 2/18 = 4/9, 1/36 + 0 = 1/36) along either axis.  Two independent pins of the composition."""
    n_u, ny, n_p, steps = 4, 24, 28, 200          # uniform extent, wall-normal extent, periodic in-plane extent
    p2 = P.sc_params(P.MODEL_SC_D2Q9, n_p, ny, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT)
    o2 = OracleSim(p2).init_case(P.CASE_SC_CONTACT2D, (0.265, 0.038, 6.0))
    rho2d = o2.fields()["s0"].reshape(n_p, ny)                  # [x2][y]
    fl2d = o2.flag.reshape(n_p, ny)
    T19 = np.array([1 / 18.] * 3 + [1 / 36.] * 6 + [1 / 3.] + [1 / 18.] * 3 + [1 / 36.] * 6)
    if axis == "z":      # 3-D x = 2-D x, uniform in z
        shape = (n_p, ny, n_u)
        rho3d = np.repeat(rho2d[:, :, None], n_u, axis=2)
        fl3d = np.repeat(fl2d[:, :, None], n_u, axis=2)
        back = lambda v: v.reshape(shape)[:, :, 0].reshape(-1)
        pairs, zero = (("s0", "s0"), ("ux", "ux"), ("uy", "uy")), "uz"
    else:                # 3-D z = 2-D x, uniform in x
        shape = (n_u, ny, n_p)
        rho3d = np.repeat(rho2d.T[None, :, :], n_u, axis=0)
        fl3d = np.repeat(fl2d.T[None, :, :], n_u, axis=0)
        back = lambda v: v.reshape(shape)[0].T.reshape(-1)
        pairs, zero = (("s0", "s0"), ("uz", "ux"), ("uy", "uy")), "ux"
    p3 = P.sc_params(P.MODEL_SC_D3Q19, *shape, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT)
    o3 = OracleSim(p3)
    o3.lattice[:19 * p3.nelem] = (T19[:, None] * rho3d.reshape(-1)[None, :]).reshape(-1)
    o3.flag[:] = fl3d.reshape(-1)
    o3.step(steps)
    o2.step(steps)
    got, ref = o3.fields(), o2.fields()
    for k3, k2 in pairs:
        assert rel_linf(back(got[k3]), ref[k2]) < 1e-12, (axis, k3)
    assert np.max(np.abs(got[zero])) < 1e-14
    assert np.max(np.abs(ref["ux"])) > 1e-6
