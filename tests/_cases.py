"""Shared case definitions for the tests: golden fixture -> (Params, init case, steps)."""
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_package():
    """import the hyphenated package directory multiphase-lbm_b200/ as `multiphase_lbm_b200`."""
    name = "multiphase_lbm_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = os.path.join(ROOT, "multiphase-lbm_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


pkg = load_package()
P = pkg.params


def golden_names(prefix="", long_horizon=False):
    """multiphase fixtures (clbm_oracle.c); the Pulsatile fixtures have their own loader (pulsatile_golden_names).
    `*_long` fixtures (hundreds to 1000 steps of the untouched functor) pin the ORACLE in the CPU suite only
    (long_horizon=True); the device meets that horizon against the oracle in test_gpu_parity.py."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN)
                  if f.endswith(".npz") and f.startswith(prefix) and not f.startswith(("pulsatile_", "yl2d_"))
                  and (long_horizon or not f.endswith("_long.npz")))


def pulsatile_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith("pulsatile_"))


def yl2d_golden_names(long_horizon=False):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith("yl2d_")
                  and (long_horizon or not f.endswith("_long.npz")))


def load_yl2d_golden(name):
    """-> (npz, keyword arguments of the run incl. nx, ny, steps)"""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, json.loads(bytes(z["params"]).decode())


def load_pulsatile_golden(name):
    """-> (npz, N, dump steps, keyword arguments of the run)"""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["params"]).decode())
    return z, meta["N"], meta["dumps"], meta["kw"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    kw = json.loads(bytes(z["params"]).decode())
    return z, kw


def golden_setup(name):
    """-> (params, case_id, case_args, steps, field map golden-name -> clbm field slot)"""
    z, kw = load_golden(name)
    nx, ny, nz = kw["nx"], kw["ny"], kw.get("nz", 1)
    if name.startswith("sc_laplace2d"):
        p = P.sc_params(P.MODEL_SC_D2Q9, nx, ny, omega=kw["omega"], rho_w=kw["rho_w"], a=kw["a"], b=kw["b"], R=kw["R"],
                        TT0=kw["TT0"], gravity=kw["gravity"], sc_force=P.SC_FORCE_LAPLACE)
        return p, P.CASE_SC_LAPLACE2D, (kw["rhol"], kw["rhog"], 10.0), kw["steps"], \
            {"rho": "s0", "pressure": "s1", "ux": "ux", "uy": "uy"}
    if name.startswith("sc_contact2d"):
        p = P.sc_params(P.MODEL_SC_D2Q9, nx, ny, omega=kw["omega"], rho_w=kw["rho_w"], a=kw["a"], b=kw["b"], R=kw["R"],
                        TT0=kw["TT0"], gravity=0.0, sc_force=P.SC_FORCE_CONTACT)
        return p, P.CASE_SC_CONTACT2D, (kw["rhol"], kw["rhog"], kw["RR"]), kw["steps"], \
            {"rho": "s0", "pressure": "s1", "ux": "ux", "uy": "uy"}
    if name.startswith("sc_layered2d"):
        p = P.sc_layered_params(nx, ny, omega=kw["omega"], rhol=kw["rhol"], rhog=kw["rhog"], rho_w=kw["rho_w"], a=kw["a"], b=kw["b"],
                                R=kw["R"], TT0=kw["TT0"], gx=kw["gx"], gy=kw["gy"], G=kw["G"])
        return p, P.CASE_SC_LAYERED2D, (kw["rhol"], kw["rhog"], kw["h_lower"], float(kw["w_int"])), kw["steps"], \
            {"rho": "s0", "pressure": "s1", "ux": "ux", "uy": "uy"}
    if name.startswith("sc_rt2d"):
        p = P.sc_rt_params(nx, ny, omega=kw["omega"], g=kw["g"], gravity=kw["gravity"], rho_w=kw["rhow"], a=kw["a"], b=kw["b"])
        return p, P.CASE_SC_RT2D, (kw["rhol"], kw["rhog"]), kw["steps"], \
            {"rho": "s0", "pressure": "s1", "ux": "ux", "uy": "uy"}
    if name.startswith("hcz_layered2d"):
        p = P.hcz_layered_params(nx, ny, omega=kw["omega"], phi_l=kw["phi_l"], phi_g=kw["phi_g"], rho_l=kw["rho_l"], rho_g=kw["rho_g"],
                                 a=kw["a"], b=kw["b"], kappa=kw["kappa"], gx=kw["gx"], gx_const=kw["gx_const"])
        return p, P.CASE_HCZ_LAYERED2D, (kw["h_lower"], float(kw["w_int"])), kw["steps"], \
            {"phi": "s0", "P": "s1", "rho": "s2", "ux": "ux", "uy": "uy"}
    if name.startswith("hcz_rt2d"):
        p = P.hcz_params(P.MODEL_HCZ_D2Q9, nx, ny, omega=kw["omega"], phi_l=kw["phi_l"], phi_g=kw["phi_g"],
                         rho_l=kw["rho_l"], rho_g=kw["rho_g"], a=kw["a"], b=kw["b"], kappa=kw["kappa"], gravity=kw["gravity"])
        return p, P.CASE_HCZ_RT2D, (), kw["steps"], {"phi": "s0", "P": "s1", "rho": "s2", "ux": "ux", "uy": "uy"}
    if name.startswith("hcz_laplace3d"):
        p = P.hcz_params(P.MODEL_HCZ_D3Q19, nx, ny, nz, omega=kw["omega"], phi_l=kw["phi_l"], phi_g=kw["phi_g"],
                         rho_l=kw["rho_l"], rho_g=kw["rho_g"], a=kw["a"], b=kw["b"], kappa=kw["kappa"], gravity=kw["gravity"])
        return p, P.CASE_HCZ_LAPLACE3D, (), kw["steps"], \
            {"phi": "s0", "P": "s1", "rho": "s2", "ux": "ux", "uy": "uy", "uz": "uz"}
    raise KeyError(name)


def rel_linf(a, b):
    """relative L-inf with global-max normalisation: max|a-b| / max|b|"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    num = np.max(np.abs(a - b))
    return num / den if den > 0 else num


def rel_linf_vec(a_comps, b_comps):
    """relative L-inf of a VECTOR field: max over components of |a-b|, normalised by the largest |b| over ALL components.
    Used where one component decays to cancellation level (the x force / x velocity of the Shan-Chen Rayleigh-Taylor case
    drop to 1e-6 while y stays O(0.1)): a per-component norm would then measure round-off against round-off."""
    den = max(float(np.max(np.abs(np.asarray(b)))) for b in b_comps)
    num = max(float(np.max(np.abs(np.asarray(a) - np.asarray(b)))) for a, b in zip(a_comps, b_comps))
    return num / den if den > 0 else num
