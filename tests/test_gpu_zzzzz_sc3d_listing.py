"""The headline kernel against the reference's own 3-D Shan-Chen code, with no oracle in between.

The reference has no C++ D3Q19 Shan-Chen functor, but it ships a complete D3Q19 Shan-Chen program as a Fortran listing
(SC/apps/fortran: main.for, streamcollision.for, force.for).  tests/test_sc3d_oracle_symmetry.py restates that listing in numpy
(its own direction ordering, weights and operation order) and pins the CPU oracle to it; here the same restatement checks the
CUDA path directly, through the C ABI: a periodic lattice without solid nodes (the listing reflects on the node, the C++ case
files half-way, so walls are not comparable step for step -- they are pinned by the z- / x-uniform projections onto
contactAngle2D.h in test_gpu_parity.py).  The listing iterates collide(stream(.)), the case files stream(collide(.)): started
from S(ff_0), the device after n steps must hold S(ff_n).  Bar: 1e-10 (BASELINE.json north_star)."""
import numpy as np
import pytest

import _cases
from _oracle import OracleSim
from test_sc3d_oracle_symmetry import C19, FTK, FXC, FYC, FZC, _fortran_iteration, _fortran_stream

pytestmark = pytest.mark.gpu

pkg = _cases.pkg
P = pkg.params
TOL = 1e-10


@pytest.mark.parametrize("fused", [1, 0])          # the default (TMA-staged) kernel and the staged path
def test_sc_d3q19_device_equals_the_reference_fortran_listing(fused):
    nx, ny, nz, steps, tau = 24, 28, 32, 200, 1.0
    p = P.sc_params(P.MODEL_SC_D3Q19, nx, ny, nz, tau=tau, sc_force=P.SC_FORCE_LAPLACE)
    x, y, z = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    r = np.sqrt((x - 11.4) ** 2 + (y - 13.2) ** 2 + (z - 15.7) ** 2)
    rho0 = 0.1515 - 0.1135 * np.tanh((r - 9.0) / 1.5) + 0.003 * np.cos(0.7 * x - 0.4 * y + 1.1 * z)
    ff = FTK[:, None, None, None] * rho0[None]
    to19 = [int(np.where((C19 == (FXC[k], FYC[k], FZC[k])).all(axis=1))[0][0]) for k in range(19)]     # listing k -> laplace3D.h k

    def layout(ff_post):                              # S(ff) in the case files' ordering, [19][nelem]
        s = _fortran_stream(ff_post)
        out = np.empty((19, nx * ny * nz))
        for k in range(19):
            out[to19[k]] = s[k].reshape(-1)
        return out

    host = OracleSim(p)                               # only as the holder of a reference-layout host state; it never steps
    host.lattice[:19 * p.nelem] = layout(ff).reshape(-1)
    with pkg.clbm.Lattice(p.copy(fused=fused)) as lat:
        lat.upload(host.lattice, host.flag, 0)
        lat.step(steps)
        got = lat.fields()
        pops = lat.in_pops()[0]
    for _ in range(steps):
        ff, rho, u, F = _fortran_iteration(ff, tau, p.TT)
    assert _cases.rel_linf(pops, layout(ff)) < TOL
    _, rho, u, F = _fortran_iteration(ff, tau, p.TT)  # the listing's next pass: moments and force of S(ff_n)
    up = u + F / 2.0 / rho                            # calcu_upr, the "real fluid velocity" (= u_actual of the case files)
    assert rho.max() - rho.min() > 0.15               # still a droplet
    assert _cases.rel_linf(got["s0"], rho.reshape(-1)) < TOL
    vel = np.stack([got["ux"], got["uy"], got["uz"]])
    ref = up.reshape(3, -1)
    assert np.max(np.abs(ref)) > 1e-6
    assert np.max(np.abs(vel - ref)) / np.max(np.abs(ref)) < TOL      # the velocity as a vector (global-max normalisation)
