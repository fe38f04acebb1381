#!/usr/bin/env python
"""BASELINE.json configs[0] and configs[1] IN FULL through the UNTOUCHED reference functors (run where /root/reference exists):

    make -C oracle ref && python tests/golden/make_golden_baseline_configs.py

configs[0]: Shan-Chen D2Q9 static droplet 256 x 256, config_Laplace2D.txt parameters (omega from ulb = .01, Re = 6), 1000 steps;
configs[1]: HCZ D2Q9 Rayleigh-Taylor 256 x 1026, config_rayleighTaylor2D.txt parameters, 1000 steps (two minutes of the functor
on 8 threads).  The arrays are 5 - 40 MB, so what is committed is the SHA-256 of the populations, every field and the mask plus a
few readable scalars -> tests/golden/baseline_configs_sha256.json; tests/test_oracle_vs_reference.py replays both through the
oracle.  The GPU suite compares the device with the oracle on exactly these two configurations (test_gpu_parity.py), which
closes the chain reference == oracle (bit for bit) ~ device (1e-10) at the two BASELINE configurations a CPU can finish."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
P = entry.load_package().params
THREADS = max(1, len(os.sched_getaffinity(0)))

OM_C1 = P.lb_parameters(0.01, 256, 6.0)[1]
OM_C2 = P.lb_parameters(0.04, 256, 3000.0)[1]
OM_C3 = P.lb_parameters(0.04, 2048, 3000.0)[1]
CASES = {
    "c1_sc_d2q9_256": ("ref_sc_laplace2d", dict(nx=256, ny=256, steps=1000, omega=OM_C1, rhol=0.265, rhog=0.038, rho_w=0.12,
                                               a=1.0, b=4.0, R=1.0, TT0=0.875, gravity=0.0), 1, 9, ["rho", "pressure", "ux", "uy"]),
    "c2_hcz_d2q9_256": ("ref_hcz_rt2d", dict(nx=256, ny=1026, steps=1000, omega=OM_C2, phi_l=0.251, phi_g=0.024, rho_l=0.12,
                                            rho_g=0.04, a=4.0, b=4.0, kappa=0.01, gravity=-6.25e-6), 2, 9, ["phi", "P", "rho", "ux", "uy"]),
    # configs[2] (2048 x 8194) is 36 h of the functor: its COLUMN shape and parameters on 16 columns, 100 steps, as the GPU suite runs it
    "c3_shape_16x8194": ("ref_hcz_rt2d", dict(nx=16, ny=8194, steps=100, omega=OM_C3, phi_l=0.251, phi_g=0.024, rho_l=0.12,
                                             rho_g=0.04, a=4.0, b=4.0, kappa=0.01, gravity=-6.25e-6), 2, 9, ["phi", "P", "rho", "ux", "uy"]),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    path = os.path.join(HERE, "baseline_configs_sha256.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for name, (binary, kw, sets, Q, fields) in CASES.items():
        if sys.argv[1:] and name not in sys.argv[1:]:
            continue
        with tempfile.TemporaryDirectory() as td:
            dump = os.path.join(td, "dump.bin")
            log = subprocess.check_output([os.path.join(REF, binary)] + ["%s=%r" % kv for kv in kw.items()] +
                                          ["threads=%d" % THREADS, "out=" + dump]).decode().strip()
            print(name, log)
            raw = np.fromfile(dump, dtype=np.uint8)
        ne = kw["nx"] * kw["ny"]
        nd = sets * Q * ne + len(fields) * ne
        assert raw.size == nd * 8 + ne
        dbl = raw[:nd * 8].view(np.float64)
        rec = {"binary": binary, "params": kw, "sets": sets, "Q": Q, "fields": fields, "sha256": {}, "max_abs": {}}
        rec["sha256"]["pops"] = sha(dbl[:sets * Q * ne])
        off = sets * Q * ne
        for f in fields:
            rec["sha256"][f] = sha(dbl[off:off + ne])
            rec["max_abs"][f] = float(np.abs(dbl[off:off + ne]).max())
            off += ne
        rec["sha256"]["flag"] = sha(raw[nd * 8:])
        rec["bulk_nodes"] = int((raw[nd * 8:] == 1).sum())
        out[name] = rec
    json.dump(out, open(os.path.join(HERE, "baseline_configs_sha256.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
