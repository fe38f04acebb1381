#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNTOUCHED reference (oracle/_ref harness binaries).

Run in the build container (where /root/reference exists):
    make -C oracle ref && python tests/golden/make_golden.py
Each fixture holds the reference's "in" populations and macroscopic fields after `steps`
steps of the reference functor from the reference's own initial condition, plus the exact
parameters used, so tests can replay the same case through the oracle and the CUDA path.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")

# name -> (binary, kwargs, sets, Q, field names)
CASES = {
    "sc_laplace2d_32x32_s200": ("ref_sc_laplace2d", dict(nx=32, ny=32, steps=200, omega=0.5617977528089888,
                                rhol=0.265, rhog=0.038, rho_w=0.12, a=1.0, b=4.0, R=1.0, TT0=0.875, gravity=0.0),
                                1, 9, ["rho", "pressure", "ux", "uy"]),
    "sc_laplace2d_grav_24x40_s60": ("ref_sc_laplace2d", dict(nx=24, ny=40, steps=60, omega=1.2, rhol=0.265, rhog=0.038,
                                    rho_w=0.12, a=1.0, b=4.0, R=1.0, TT0=0.875, gravity=-1e-5),
                                    1, 9, ["rho", "pressure", "ux", "uy"]),
    "sc_contact2d_48x24_s200": ("ref_sc_contact2d", dict(nx=48, ny=24, steps=200, omega=1.0, rhol=0.265, rhog=0.038,
                                rho_w=0.2, a=1.0, b=4.0, R=1.0, TT0=0.875, RR=8.0),
                                1, 9, ["rho", "pressure", "ux", "uy"]),
    "sc_layered2d_10x41_s300": ("ref_sc_layered2d", dict(nx=10, ny=41, steps=300, omega=1.0, rhol=0.21, rhog=0.067, rho_w=0.067,
                                a=1.0, b=4.0, R=1.0, TT0=0.95, gx=1e-6, gy=0.0, G=-1.0, h_lower=0.3, w_int=4),
                                1, 9, ["rho", "pressure", "ux", "uy"]),
    "sc_layered2d_12x33_s200": ("ref_sc_layered2d", dict(nx=12, ny=33, steps=200, omega=1.25, rhol=0.247, rhog=0.0405, rho_w=0.06,
                                a=1.0, b=4.0, R=1.0, TT0=0.9, gx=2e-6, gy=-1e-7, G=-1.3, h_lower=0.25, w_int=2),
                                1, 9, ["rho", "pressure", "ux", "uy"]),
    "sc_rt2d_16x66_s300": ("ref_sc_rt2d", dict(nx=16, ny=66, steps=300, omega=1.0, rhol=1.2, rhog=0.4, rhow=0.2, g=-5.0, a=1.0, b=4.0,
                           gravity=-1.25e-5), 1, 9, ["rho", "pressure", "ux", "uy", "fx", "fy"]),
    "sc_rt2d_20x42_s150": ("ref_sc_rt2d", dict(nx=20, ny=42, steps=150, omega=1.3, rhol=1.5, rhog=0.3, rhow=0.2, g=-4.5, a=1.0, b=4.0,
                           gravity=-5e-5), 1, 9, ["rho", "pressure", "ux", "uy", "fx", "fy"]),
    "hcz_rt2d_16x66_s40": ("ref_hcz_rt2d", dict(nx=16, ny=66, steps=40, omega=1.9598595172738, phi_l=0.251, phi_g=0.024,
                           rho_l=0.12, rho_g=0.04, a=4.0, b=4.0, kappa=0.01, gravity=-6.25e-6),
                           2, 9, ["phi", "P", "rho", "ux", "uy"]),
    "hcz_rt2d_24x50_s25": ("ref_hcz_rt2d", dict(nx=24, ny=50, steps=25, omega=1.7, phi_l=0.251, phi_g=0.024,
                           rho_l=0.12, rho_g=0.04, a=4.0, b=4.0, kappa=0.01, gravity=-2e-5),
                           2, 9, ["phi", "P", "rho", "ux", "uy"]),
    "hcz_layered2d_10x41_s60": ("ref_hcz_layered2d", dict(nx=10, ny=41, steps=60, omega=1.0, phi_l=0.251, phi_g=0.024, rho_l=0.12,
                                rho_g=0.04, a=4.0, b=4.0, kappa=0.001, gx=0.0, gx_const=1e-6, h_lower=0.3, w_int=2),
                                2, 9, ["phi", "P", "rho", "ux", "uy"]),
    "hcz_layered2d_12x29_s45": ("ref_hcz_layered2d", dict(nx=12, ny=29, steps=45, omega=1.4, phi_l=0.251, phi_g=0.024, rho_l=0.12,
                                rho_g=0.04, a=4.0, b=4.0, kappa=0.004, gx=2e-6, gx_const=5e-7, h_lower=0.25, w_int=3),
                                2, 9, ["phi", "P", "rho", "ux", "uy"]),
    "hcz_laplace3d_8x8x8_s4": ("ref_hcz_laplace3d", dict(nx=8, ny=8, nz=8, steps=4, omega=0.5617977528089888,
                               phi_l=0.251, phi_g=0.024, rho_l=0.12, rho_g=0.04, a=4.0, b=4.0, kappa=5e-4, gravity=0.0),
                               2, 19, ["phi", "P", "rho", "ux", "uy", "uz"]),
    # the north_star horizon against the untouched functor itself (CPU suite only: tests/test_oracle_vs_reference.py)
    "sc_laplace2d_32x32_s1000_long": ("ref_sc_laplace2d", dict(nx=32, ny=32, steps=1000, omega=0.5617977528089888,
                                      rhol=0.265, rhog=0.038, rho_w=0.12, a=1.0, b=4.0, R=1.0, TT0=0.875, gravity=0.0),
                                      1, 9, ["rho", "pressure", "ux", "uy"]),
    "sc_contact2d_48x24_s1000_long": ("ref_sc_contact2d", dict(nx=48, ny=24, steps=1000, omega=1.0, rhol=0.265, rhog=0.038,
                                      rho_w=0.2, a=1.0, b=4.0, R=1.0, TT0=0.875, RR=8.0),
                                      1, 9, ["rho", "pressure", "ux", "uy"]),
    "sc_layered2d_10x41_s1000_long": ("ref_sc_layered2d", dict(nx=10, ny=41, steps=1000, omega=1.0, rhol=0.21, rhog=0.067, rho_w=0.067,
                                      a=1.0, b=4.0, R=1.0, TT0=0.95, gx=1e-6, gy=0.0, G=-1.0, h_lower=0.3, w_int=4),
                                      1, 9, ["rho", "pressure", "ux", "uy"]),
    "sc_rt2d_16x66_s1000_long": ("ref_sc_rt2d", dict(nx=16, ny=66, steps=1000, omega=1.0, rhol=1.2, rhog=0.4, rhow=0.2, g=-5.0, a=1.0, b=4.0,
                                 gravity=-1.25e-5), 1, 9, ["rho", "pressure", "ux", "uy", "fx", "fy"]),
    "hcz_layered2d_10x41_s1000_long": ("ref_hcz_layered2d", dict(nx=10, ny=41, steps=1000, omega=1.0, phi_l=0.251, phi_g=0.024, rho_l=0.12,
                                       rho_g=0.04, a=4.0, b=4.0, kappa=0.001, gx=0.0, gx_const=1e-6, h_lower=0.3, w_int=2),
                                       2, 9, ["phi", "P", "rho", "ux", "uy"]),
    "hcz_rt2d_16x66_s1000_long": ("ref_hcz_rt2d", dict(nx=16, ny=66, steps=1000, omega=1.9598595172738, phi_l=0.251, phi_g=0.024,
                                  rho_l=0.12, rho_g=0.04, a=4.0, b=4.0, kappa=0.01, gravity=-6.25e-6),
                                  2, 9, ["phi", "P", "rho", "ux", "uy"]),
    "hcz_laplace3d_10x8x12_s400_long": ("ref_hcz_laplace3d", dict(nx=10, ny=8, nz=12, steps=400, omega=1.3,
                                        phi_l=0.251, phi_g=0.024, rho_l=0.12, rho_g=0.04, a=4.0, b=4.0, kappa=5e-4, gravity=-1e-5),
                                        2, 19, ["phi", "P", "rho", "ux", "uy", "uz"]),
    "hcz_laplace3d_10x6x12_s3": ("ref_hcz_laplace3d", dict(nx=10, ny=6, nz=12, steps=3, omega=1.3,
                                 phi_l=0.251, phi_g=0.024, rho_l=0.12, rho_g=0.04, a=4.0, b=4.0, kappa=5e-4, gravity=-1e-5),
                                 2, 19, ["phi", "P", "rho", "ux", "uy", "uz"]),
}


def run_case(name):
    binary, kw, sets, Q, fields = CASES[name]
    exe = os.path.join(REF, binary)
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "dump.bin")
        args = [exe] + ["%s=%r" % (k, v) for k, v in kw.items()] + ["out=" + out]
        log = subprocess.check_output(args).decode()
        raw = np.fromfile(out, dtype=np.uint8)
    ne = kw["nx"] * kw["ny"] * kw.get("nz", 1)
    nd = sets * Q * ne + len(fields) * ne
    dbl = raw[:nd * 8].view(np.float64)
    flag = raw[nd * 8:nd * 8 + ne].copy()
    assert raw.size == nd * 8 + ne, (raw.size, nd * 8 + ne)
    data = {"pops": dbl[:sets * Q * ne].reshape(sets, Q, ne).copy(), "flag": flag,
            "params": np.frombuffer(json.dumps(kw).encode(), dtype=np.uint8)}
    off = sets * Q * ne
    for f in fields:
        data[f] = dbl[off:off + ne].copy()
        off += ne
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
    print(name, log.strip())


if __name__ == "__main__":
    for n in (sys.argv[1:] or CASES):
        run_case(n)
