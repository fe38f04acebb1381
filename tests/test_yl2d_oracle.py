"""oracle/yl2d_oracle.c must reproduce the reference's Young-Laplace case (AB/apps/Young_Laplace2D.h, the AB build's
default problem) bit-for-bit: fields C, P, Rho, Ux, Uy, all four population buffers and the parity, against binary dumps
of the untouched header (tests/golden/yl2d_*.npz, made by tests/golden/make_golden_yl2d.py)."""
import numpy as np
import pytest

import _cases
from _oracle import YL2DOracle


@pytest.mark.parametrize("name", _cases.yl2d_golden_names(long_horizon=True))
def test_yl2d_oracle_bit_exact_vs_reference(name):
    z, kw = _cases.load_yl2d_golden(name)
    steps = kw.pop("steps")
    o = YL2DOracle(**kw).step(steps)
    f = o.fields()
    for k in ("C", "P", "Rho", "Ux", "Uy"):
        np.testing.assert_array_equal(f[k], z[k], err_msg=k)
    np.testing.assert_array_equal(o.lattice(), z["lattice"])
    assert o.parity == int(z["parity"])
    o.close()


def test_yl2d_mass_is_conserved():
    o = YL2DOracle(48, 48)
    m0 = o.fields()["C"].sum()
    o.step(300)
    assert abs(o.fields()["C"].sum() - m0) < 1e-11 * m0
    o.close()
