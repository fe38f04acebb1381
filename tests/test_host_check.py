"""Per-cell Shan-Chen device functions (csrc/sc_cell.cuh) compiled for the HOST and checked against the oracle.

Test infrastructure for containers without a GPU (tests/host_check/sc_cell_host.cu): the same psi / force / collision /
output functions the kernels inline, driven by the two loops of the staged kernels.  It does not replace the `-m gpu`
parity tests (kernel indexing, shared-memory staging, fused pipelines); it catches arithmetic slips in a variant before
GPU time is spent on it.  Tolerance: 1e-10 relative L-inf (global-max normalisation), the bar of the GPU tests.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import _cases
from _cases import P, rel_linf, rel_linf_vec
from _oracle import OracleSim

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_check", "sc_cell_host.cu")
LIB = os.path.join(HERE, "host_check", "_build", "libsc_cell_host.so")
CSRC = os.path.join(_cases.ROOT, "multiphase-lbm_b200", "csrc")
TOL = 1e-10

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists(LIB), reason="nvcc not available")


def _lib():
    deps = [SRC, os.path.join(_cases.ROOT, "include", "clbm.h")] + \
           [os.path.join(CSRC, f) for f in ("sc_cell.cuh", "moments.cuh", "lattice.cuh", "clbm_internal.h", "mrt.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["nvcc", "-std=c++17", "-O2", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared", "-o", LIB, SRC])
    L = ctypes.CDLL(LIB)
    L.host_check_sc_step.restype = ctypes.c_int
    L.host_check_sc_fields.restype = ctypes.c_int
    return L


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class HostSim:
    def __init__(self, osim):
        self.p = osim.p
        self.lattice = osim.lattice.copy()
        self.flag = osim.flag.copy()
        self.parity = ctypes.c_int(osim.parity.value)

    def step(self, n):
        rc = _lib().host_check_sc_step(ctypes.byref(self.p), _dp(self.lattice), self.flag.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                       ctypes.byref(self.parity), int(n))
        assert rc == 0
        return self

    def in_pops(self):
        npop = self.p.Q * self.p.nelem
        off = self.parity.value * npop
        return self.lattice[off:off + npop]

    def fields(self):
        names = ["s0", "s1", "ux", "uy", "uz", "fx", "fy", "fz"]
        arrs = {k: np.zeros(self.p.nelem) for k in names}
        ptrs = (ctypes.POINTER(ctypes.c_double) * 8)(*[_dp(arrs[k]) for k in names])
        rc = _lib().host_check_sc_fields(ctypes.byref(self.p), _dp(self.lattice), self.flag.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                         self.parity.value, ptrs)
        assert rc == 0
        return arrs


def _compare(p, case_id, args, steps, vec_norm=False):
    o = OracleSim(p).init_case(case_id, args)
    h = HostSim(o)
    o.step(steps)
    h.step(steps)
    assert h.parity.value == o.parity.value
    npop = p.Q * p.nelem
    ref = o.lattice[o.parity.value * npop:(o.parity.value + 1) * npop]
    assert rel_linf(h.in_pops(), ref) < TOL
    fo, fh = o.fields(), h.fields()
    for k in ("s0", "s1"):
        assert rel_linf(fh[k], fo[k]) < TOL, k
    Fo = o.force()
    if vec_norm:
        assert rel_linf_vec([fh[k] for k in ("ux", "uy", "uz")], [fo[k] for k in ("ux", "uy", "uz")]) < TOL
        assert rel_linf_vec([fh[k] for k in ("fx", "fy", "fz")], [Fo[k] for k in ("fx", "fy", "fz")]) < TOL
    else:
        for k in ("ux", "uy", "uz"):
            assert rel_linf(fh[k], fo[k]) < TOL, k
        for k in ("fx", "fy", "fz"):
            assert rel_linf(fh[k], Fo[k]) < TOL, k
    return fo


@pytest.mark.parametrize("name", _cases.golden_names("sc_"))
def test_host_cell_functions_vs_oracle_on_golden_cases(name):
    """every Shan-Chen fixture (Laplace, contact angle, constant-G layered, Rayleigh-Taylor/Guo): same case, same step count"""
    p, case_id, args, steps, _ = _cases.golden_setup(name)
    _compare(p, case_id, args, steps)


def test_host_cell_functions_sc_rt_1000_steps():
    """SC/apps/RayleighTaylor2D.h at the shipped parameters (omega = 1, g = -5, gravity = -1.25e-5), 1000 steps"""
    p = P.sc_rt_params(24, 98, omega=1.0)
    f = _compare(p, P.CASE_SC_RT2D, (1.2, 0.4), 1000, vec_norm=True)   # x components decay to 1e-6: see rel_linf_vec
    assert np.max(np.abs(f["uy"])) > 1e-6


def test_host_cell_functions_sc_d3q19_walls():
    p = P.sc_params(P.MODEL_SC_D3Q19, 12, 10, 8, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT)
    _compare(p, P.CASE_SC_DROPLET3D, (0.265, 0.038, 3.0, 4.0), 50)


def test_mrt9_factorised_form_equals_the_matrix_form():
    """csrc/mrt.cuh against M^-1 S M built from the moment rows of CooLBM_MRT_combustion.cpp:313-323 in the k-ordering of the
    case headers; S = omega I must be omega times the identity"""
    c = np.array([(-1, 0), (0, -1), (-1, -1), (-1, 1), (0, 0), (1, 0), (0, 1), (1, 1), (1, -1)], dtype=np.float64)
    cx, cy = c[:, 0], c[:, 1]
    c2 = cx * cx + cy * cy
    M = np.stack([np.ones(9), -4 + 3 * c2, 4 - 10.5 * c2 + 4.5 * c2 * c2, cx, (-5 + 3 * c2) * cx, cy, (-5 + 3 * c2) * cy,
                  cx * cx - cy * cy, cx * cy])
    # the same matrix as the reference's table (rest first, E, N, W, S, NE, NW, SW, SE), permuted
    ref_order = [(0, 0), (1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1)]
    Mref = np.array([[1, 1, 1, 1, 1, 1, 1, 1, 1], [-4, -1, -1, -1, -1, 2, 2, 2, 2], [4, -2, -2, -2, -2, 1, 1, 1, 1],
                     [0, 1, 0, -1, 0, 1, -1, -1, 1], [0, -2, 0, 2, 0, 1, -1, -1, 1], [0, 0, 1, 0, -1, 1, 1, -1, -1],
                     [0, 0, -2, 0, 2, 1, 1, -1, -1], [0, 1, -1, 1, -1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 1, -1, 1, -1]], dtype=np.float64)
    perm = [ref_order.index((int(a), int(b))) for a, b in c]
    np.testing.assert_array_equal(M, Mref[:, perm])
    Minv = np.linalg.inv(M)
    # ... and its inverse is the reference's literal M_inv table (CooLBM_MRT_combustion.cpp:326-336), rows permuted the same way
    Minv_ref = np.array([[1 / 9., -1 / 9., 1 / 9., 0, 0, 0, 0, 0, 0],
                         [1 / 9., -1 / 36., -1 / 18., 1 / 6., -1 / 6., 0, 0, 1 / 4., 0],
                         [1 / 9., -1 / 36., -1 / 18., 0, 0, 1 / 6., -1 / 6., -1 / 4., 0],
                         [1 / 9., -1 / 36., -1 / 18., -1 / 6., 1 / 6., 0, 0, 1 / 4., 0],
                         [1 / 9., -1 / 36., -1 / 18., 0, 0, -1 / 6., 1 / 6., -1 / 4., 0],
                         [1 / 9., 1 / 18., 1 / 36., 1 / 6., 1 / 12., 1 / 6., 1 / 12., 0, 1 / 4.],
                         [1 / 9., 1 / 18., 1 / 36., -1 / 6., -1 / 12., 1 / 6., 1 / 12., 0, -1 / 4.],
                         [1 / 9., 1 / 18., 1 / 36., -1 / 6., -1 / 12., -1 / 6., -1 / 12., 0, 1 / 4.],
                         [1 / 9., 1 / 18., 1 / 36., 1 / 6., 1 / 12., -1 / 6., -1 / 12., 0, -1 / 4.]])
    np.testing.assert_allclose(Minv_ref @ Mref, np.eye(9), atol=2e-16)
    np.testing.assert_allclose(Minv, Minv_ref[perm, :], atol=2e-16)
    L = _lib()
    rng = np.random.default_rng(1)
    for rates in ([1.3, 1.3, 1.3, 1.3, 1.3], [1.7, 1.1, 1.2, 0.9, 1.7], [1.0, 0.5, 1.9, 1.4, 1.96]):
        s_c, s_e, s_eps, s_q, s_nu = rates
        S = np.diag([s_c, s_e, s_eps, s_c, s_q, s_c, s_q, s_nu, s_nu])
        A = Minv @ S @ M
        for _ in range(20):
            v = rng.standard_normal(9)
            w = np.zeros(9)
            L.host_check_mrt9(_dp(v), _dp(np.asarray(rates, dtype=np.float64)), _dp(w))
            assert np.max(np.abs(w - A @ v)) < 5e-15 * np.max(np.abs(v))
            if len(set(rates)) == 1:
                assert np.max(np.abs(w - rates[0] * v)) < 5e-15 * np.max(np.abs(v))


@pytest.mark.parametrize("case", ["laplace", "contact", "layered"])
def test_host_cell_functions_sc_mrt(case):
    """CLBM_COLLISION_MRT for Yuan-CS Shan-Chen D2Q9: the device's sc_collide_mrt (factorised M^-1 S M) against the oracle's
    matrix form at free rates, and the oracle at S = omega I against its own BGK branch (pinned to the reference)"""
    if case == "laplace":
        mk = lambda **k: P.sc_params(P.MODEL_SC_D2Q9, 48, 40, omega=1.2, gravity=-1e-5, **k)
        cid, args, steps = P.CASE_SC_LAPLACE2D, (0.265, 0.038, 10.0), 300
    elif case == "contact":
        mk = lambda **k: P.sc_params(P.MODEL_SC_D2Q9, 48, 24, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT, **k)
        cid, args, steps = P.CASE_SC_CONTACT2D, (0.265, 0.038, 8.0), 300
    else:
        mk = lambda **k: P.sc_layered_params(10, 41, omega=1.1, gx=1e-6, **k)
        cid, args, steps = P.CASE_SC_LAYERED2D, (0.21, 0.067, 0.3, 4.0), 300
    bgk = mk()
    same = bgk.copy(collision=P.COLLISION_MRT, s_e=bgk.omega, s_eps=bgk.omega, s_q=bgk.omega)
    a = OracleSim(bgk).init_case(cid, args).step(steps)
    b = OracleSim(same).init_case(cid, args).step(steps)
    assert rel_linf(b.in_pops(), a.in_pops()) < 1e-12
    free = bgk.copy(collision=P.COLLISION_MRT, s_e=min(1.9, bgk.omega + 0.3), s_eps=max(0.5, bgk.omega - 0.2), s_q=1.4)
    f = _compare(free, cid, args, steps)
    assert np.isfinite(f["ux"]).all()
    c = OracleSim(free).init_case(cid, args).step(steps)
    assert rel_linf(c.in_pops(), a.in_pops()) > 1e-8        # the free rates do change the solution


def test_host_cell_functions_sc_d3q19_equal_the_reference_fortran_listing():
    """the DEVICE's per-cell Shan-Chen D3Q19 arithmetic (sc_cell.cuh compiled for the host) against the numpy restatement of the
    reference's Fortran D3Q19 listing (SC/apps/fortran) with no oracle in between -- the CPU-side twin of
    tests/test_gpu_zzzzz_sc3d_listing.py, same lattice, state and step count (periodic, no solid nodes: see there)."""
    from test_sc3d_oracle_symmetry import C19, FTK, FXC, FYC, FZC, _fortran_iteration, _fortran_stream
    nx, ny, nz, steps, tau = 24, 28, 32, 200, 1.0
    p = P.sc_params(P.MODEL_SC_D3Q19, nx, ny, nz, tau=tau, sc_force=P.SC_FORCE_LAPLACE)
    x, y, z = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    r = np.sqrt((x - 11.4) ** 2 + (y - 13.2) ** 2 + (z - 15.7) ** 2)
    rho0 = 0.1515 - 0.1135 * np.tanh((r - 9.0) / 1.5) + 0.003 * np.cos(0.7 * x - 0.4 * y + 1.1 * z)
    ff = FTK[:, None, None, None] * rho0[None]
    to19 = [int(np.where((C19 == (FXC[k], FYC[k], FZC[k])).all(axis=1))[0][0]) for k in range(19)]

    def layout(ff_post):
        s = _fortran_stream(ff_post)
        out = np.empty((19, nx * ny * nz))
        for k in range(19):
            out[to19[k]] = s[k].reshape(-1)
        return out

    host = OracleSim(p)                       # holder of the reference-layout arrays only; the oracle never steps here
    host.lattice[:19 * p.nelem] = layout(ff).reshape(-1)
    h = HostSim(host).step(steps)
    for _ in range(steps):
        ff, rho, u, F = _fortran_iteration(ff, tau, p.TT)
    assert h.parity.value == 0
    assert rel_linf(h.in_pops(), layout(ff).reshape(-1)) < TOL
    _, rho, u, F = _fortran_iteration(ff, tau, p.TT)
    up = u + F / 2.0 / rho
    fh = h.fields()
    assert rho.max() - rho.min() > 0.15
    assert rel_linf(fh["s0"], rho.reshape(-1)) < TOL
    assert rel_linf_vec([fh[k] for k in ("ux", "uy", "uz")], list(up.reshape(3, -1))) < TOL
    assert rel_linf_vec([fh[k] for k in ("fx", "fy", "fz")], list(F.reshape(3, -1))) < TOL


def test_host_cell_force_sc_d3q19_walls_equals_the_reference_fortran_listing():
    """the device's force gathering (sc_gather_force + the wall term of the contact-angle variant, host-compiled) against
    `calcu_Fxy` of the Fortran listing on a 3-D state with the two wall planes and a solid block in the bulk"""
    from test_sc3d_oracle_symmetry import T19, _fortran_calcu_Fxy
    nx, ny, nz = 14, 12, 10
    p = P.sc_params(P.MODEL_SC_D3Q19, nx, ny, nz, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT)
    o = OracleSim(p)
    x, y, z = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    r = np.sqrt((x - 6.3) ** 2 + (y - 3.1) ** 2 + (z - 4.6) ** 2)
    rho = 0.1515 - 0.1135 * np.tanh((r - 3.7) / 1.3) + 0.004 * np.sin(0.9 * x + 0.5 * y - 1.3 * z)
    solid = (y == 0) | (y == ny - 1) | ((x >= 9) & (x <= 10) & (y >= 5) & (y <= 7) & (z >= 2) & (z <= 4))
    o.flag[:] = np.where(solid, 0, 1).reshape(-1).astype(np.uint8)
    o.lattice[:19 * p.nelem] = (T19[:, None] * rho.reshape(-1)[None, :]).reshape(-1)
    fh = HostSim(o).fields()
    got = np.stack([fh["fx"], fh["fy"], fh["fz"]], axis=-1).reshape(nx, ny, nz, 3)
    ref, _ = _fortran_calcu_Fxy(rho, solid, p.rho_w, p.TT)
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-13
