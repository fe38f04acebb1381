"""ctypes binding of oracle/_build/liboracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is the checker, never the product path: only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

_lib = None


def build(force=False):
    src = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith(".c")]
    src.append(os.path.join(ROOT, "include", "clbm.h"))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.oracle_lattice_size.restype = ctypes.c_size_t
        _lib.oracle_step.restype = ctypes.c_int
        _lib.oracle_fields.restype = ctypes.c_int
        _lib.oracle_init_case.restype = ctypes.c_int
        _lib.oracle_force.restype = ctypes.c_int
    return _lib


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


class OracleSim:
    """Reference-layout host state advanced by the CPU oracle."""

    def __init__(self, params):
        self.p = params
        self.lattice = np.zeros(params.lattice_size, dtype=np.float64)
        self.flag = np.ones(params.nelem, dtype=np.uint8)
        self.parity = ctypes.c_int(0)

    def init_case(self, case_id, args=()):
        a = np.asarray(args, dtype=np.float64)
        rc = lib().oracle_init_case(ctypes.byref(self.p), int(case_id), _dptr(a), int(a.size), _dptr(self.lattice),
                                    self.flag.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), ctypes.byref(self.parity))
        assert rc == 0, "oracle_init_case failed"
        return self

    def step(self, n=1, threads=0):
        rc = lib().oracle_step(ctypes.byref(self.p), _dptr(self.lattice),
                               self.flag.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                               ctypes.byref(self.parity), int(n), int(threads))
        assert rc == 0
        return self

    def in_pops(self):
        """current ("in") populations as [sets, Q, nelem]"""
        p = self.p
        npop = p.Q * p.nelem
        out = []
        for s in range(p.sets):
            off = s * 2 * npop + self.parity.value * npop
            out.append(self.lattice[off:off + npop].reshape(p.Q, p.nelem))
        return np.stack(out)

    def fields(self):
        n = self.p.nelem
        names = ["s0", "s1", "s2", "ux", "uy", "uz"]
        arrs = {k: np.zeros(n) for k in names}
        rc = lib().oracle_fields(ctypes.byref(self.p), _dptr(self.lattice),
                                 self.flag.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), self.parity.value,
                                 *[_dptr(arrs[k]) for k in names])
        assert rc == 0
        return arrs

    def force(self):
        """interaction force of the Shan-Chen models (0 at non-bulk nodes): {"fx", "fy", "fz"}"""
        n = self.p.nelem
        arrs = {k: np.zeros(n) for k in ("fx", "fy", "fz")}
        rc = lib().oracle_force(ctypes.byref(self.p), _dptr(self.lattice),
                                self.flag.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), self.parity.value,
                                *[_dptr(arrs[k]) for k in ("fx", "fy", "fz")])
        assert rc == 0
        return arrs


def max_threads():
    return lib().oracle_max_threads()


def ref_binary(name):
    """path of a prebuilt reference-harness binary (oracle/_ref), or None"""
    path = os.path.join(REF_DIR, name)
    return path if os.path.exists(path) else None


class PulsatileOracle:
    """oracle/pulsatile_oracle.c: CPU restatement of AB/apps/PulsatileBloodFlow2D.h (test infrastructure)."""

    def __init__(self, N=64, tau=0.75, alpha=0.01, p0_in=0.20, p0_out=0.19, is_severed=1, deformable=1):
        L = lib()
        L.pulsatile_create.restype = ctypes.c_void_p
        L.pulsatile_create.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_int, ctypes.c_int]
        L.pulsatile_lattice.restype = ctypes.POINTER(ctypes.c_double)
        for f in ("pulsatile_step", "pulsatile_get", "pulsatile_destroy", "pulsatile_nx", "pulsatile_ny",
                  "pulsatile_parity", "pulsatile_tf", "pulsatile_fresh", "pulsatile_lattice"):
            getattr(L, f).argtypes = None
        self.h = ctypes.c_void_p(L.pulsatile_create(N, tau, alpha, p0_in, p0_out, is_severed, deformable))
        if not self.h:
            raise RuntimeError("Initial wall location out of bounds.")
        self.nx, self.ny = L.pulsatile_nx(self.h), L.pulsatile_ny(self.h)
        self.nelem = self.nx * self.ny
        self.tf = L.pulsatile_tf(self.h)

    def step(self, n=1):
        lib().pulsatile_step(self.h, int(n))
        return self

    def fields(self):
        ne = self.nelem
        out = {"P": np.zeros(ne), "Ux": np.zeros(ne), "Uy": np.zeros(ne), "flag": np.zeros(ne, dtype=np.uint8),
               "yr1": np.zeros(self.nx), "yr2": np.zeros(self.nx)}
        lib().pulsatile_get(self.h, _dptr(out["P"]), _dptr(out["Ux"]), _dptr(out["Uy"]),
                            out["flag"].ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), _dptr(out["yr1"]), _dptr(out["yr2"]))
        return out

    def lattice(self):
        return np.ctypeslib.as_array(lib().pulsatile_lattice(self.h), shape=(2 * 9 * self.nelem,)).copy()

    def set_state(self, lattice, P, Ux, Uy, yr1, yr2, parity=0, t_iter=0):
        L = lib()
        L.pulsatile_set_state.argtypes = None
        c = [np.ascontiguousarray(a, dtype=np.float64) for a in (lattice, P, Ux, Uy, yr1, yr2)]
        assert c[0].size == 2 * 9 * self.nelem and c[1].size == self.nelem and c[4].size == self.nx
        L.pulsatile_set_state(self.h, *[_dptr(a) for a in c], int(parity), int(t_iter))
        return self

    @property
    def parity(self):
        return lib().pulsatile_parity(self.h)

    def close(self):
        if self.h:
            lib().pulsatile_destroy(self.h)
            self.h = None


def pulsatile_write_vtk(nx, ny, P, Ux, Uy, flag, time_iter, path):
    """the reference's legacy-VTK text (AB/apps/PulsatileBloodFlow2D.h:680-706) of the given arrays"""
    L = lib()
    L.pulsatile_write_vtk.argtypes = None
    rc = L.pulsatile_write_vtk(int(nx), int(ny), _dptr(np.ascontiguousarray(P)), _dptr(np.ascontiguousarray(Ux)),
                               _dptr(np.ascontiguousarray(Uy)),
                               np.ascontiguousarray(flag).ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), int(time_iter),
                               path.encode())
    assert rc == 0


class YL2DOracle:
    """oracle/yl2d_oracle.c: CPU restatement of AB/apps/Young_Laplace2D.h (test infrastructure)."""

    def __init__(self, nx=32, ny=32, Sigma=0.01, W=4.0, M=0.02, RhoL=0.001, RhoH=1.0, tau=0.8):
        L = lib()
        L.yl2d_create.restype = ctypes.c_void_p
        L.yl2d_create.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 6
        L.yl2d_lattice.restype = ctypes.POINTER(ctypes.c_double)
        for f in ("yl2d_step", "yl2d_get", "yl2d_destroy", "yl2d_parity", "yl2d_lattice"):
            getattr(L, f).argtypes = None
        self.h = ctypes.c_void_p(L.yl2d_create(nx, ny, Sigma, W, M, RhoL, RhoH, tau))
        self.nx, self.ny, self.nelem = nx, ny, nx * ny

    def step(self, n=1):
        lib().yl2d_step(self.h, int(n))
        return self

    def fields(self):
        out = {k: np.zeros(self.nelem) for k in ("C", "P", "Rho", "Ux", "Uy")}
        lib().yl2d_get(self.h, *[_dptr(out[k]) for k in ("C", "P", "Rho", "Ux", "Uy")])
        return out

    def lattice(self):
        return np.ctypeslib.as_array(lib().yl2d_lattice(self.h), shape=(4 * 9 * self.nelem,)).copy()

    @property
    def parity(self):
        return lib().yl2d_parity(self.h)

    def close(self):
        if self.h:
            lib().yl2d_destroy(self.h)
            self.h = None
