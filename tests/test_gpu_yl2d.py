"""GPU parity of the Young-Laplace path (AB/apps/Young_Laplace2D.h) through the C ABI -- BIT-EXACT: fields C, P, Rho,
Ux, Uy and all four population buffers equal the reference dumps and the oracle after 1000 iterations.  (The model
amplifies rounding differences ~100x per 50 iterations, so the device code is built without FMA contraction in the
reference's operation order, like the Pulsatile path; a 1e-10 tolerance would not survive 1000 iterations.)"""
import numpy as np
import pytest

import _cases
from _oracle import YL2DOracle

pytestmark = pytest.mark.gpu
clbm = _cases.pkg.clbm
P = _cases.P
TOL = 1e-10


def _check(dev, ref_fields, ref_lattice, ref_parity, what=""):
    f = dev.fields()
    for k in ("C", "P", "Rho", "Ux", "Uy"):
        np.testing.assert_array_equal(f[k], ref_fields[k], err_msg="%s %s" % (what, k))
    lat, par = dev.lattice()
    assert par == ref_parity
    ne = dev.nelem
    a = lat.reshape(2, 2, 9, ne)[:, par]           # the "in" buffers of h and g
    b = np.asarray(ref_lattice).reshape(2, 2, 9, ne)[:, ref_parity]
    np.testing.assert_array_equal(a, b, err_msg=what + " populations")


@pytest.mark.parametrize("name", _cases.yl2d_golden_names())
def test_yl2d_gpu_vs_reference_dumps(name):
    z, kw = _cases.load_yl2d_golden(name)
    steps = kw.pop("steps")
    with clbm.YoungLaplace(**kw) as dev:
        dev.step(steps)
        _check(dev, z, z["lattice"], int(z["parity"]), name)


@pytest.mark.parametrize("nx,ny,steps,kw", [(128, 128, 1000, {}), (96, 64, 1000, dict(Sigma=0.02, W=5.0, M=0.05, RhoL=0.01, tau=0.7))])
def test_yl2d_gpu_vs_oracle_1000_steps(nx, ny, steps, kw):
    o = YL2DOracle(nx, ny, **kw)
    with clbm.YoungLaplace(nx, ny, **kw) as dev:
        _check(dev, o.fields(), o.lattice(), o.parity, "initial state")
        for chunk in (1, 9, steps - 10):
            o.step(chunk)
            dev.step(chunk)
        _check(dev, o.fields(), o.lattice(), o.parity, "%dx%d step %d" % (nx, ny, steps))
        f = o.fields()
        mass = dev.reduce(P.REDUCE_MASS)
        assert abs(mass - f["Rho"].sum()) < 1e-10 * abs(f["Rho"].sum())
        e_ref = 0.5 * np.sum(f["Ux"] ** 2 + f["Uy"] ** 2) / (nx * ny)
        assert abs(dev.reduce(P.REDUCE_ENERGY) - e_ref) < 1e-8 * e_ref
        assert np.max(np.abs(f["Ux"])) > 1e-9
    o.close()


def test_yl2d_upload_handover():
    """populations + the velocity update_fields produced for them, as a reference-side driver would hand them over"""
    o = YL2DOracle(48, 40).step(37)
    f = o.fields()
    with clbm.YoungLaplace(48, 40) as dev:
        dev.upload(o.lattice(), f["Ux"], f["Uy"], o.parity)
        _check(dev, f, o.lattice(), o.parity, "after upload")
        o.step(150)
        dev.step(150)
        _check(dev, o.fields(), o.lattice(), o.parity, "150 steps after upload")
    o.close()
