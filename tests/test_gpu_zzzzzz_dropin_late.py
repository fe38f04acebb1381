"""Drop-in harnesses added late in round 2 (same contract as tests/test_gpu_zu_dropin.py).

dropin_sc_contact2d: the UNTOUCHED SC/apps/contactAngle2D.h (bounce-back walls, contact-angle force: the spec source of BASELINE
configs[3]) -- density / pressure_node / u_actual and the populations at 1e-10 after 1000 steps, parity exact.

dropin_hcz_laplace3d: the HCZ D3Q19 half of north_star's metric: oracle/_ref/dropin_hcz_laplace3d #includes the UNTOUCHED
PF/apps/laplace3D.h; the REFERENCE's iniLattice + inigeom build the host state, clbm_create / clbm_upload / clbm_step /
clbm_download_lattice advance a copy, the REFERENCE's macro_phi_P / total_rho / total_P / velocity read both arrays: 1e-10,
parity exact (same contract as tests/test_gpu_zu_dropin.py for the D2Q9 headers; the reference functor runs at a few kLUPS, hence the
small lattices).  Written after the round's GPU minutes were spent: the harness logic was dry-run here against an ad-hoc
oracle-backed stand-in for libclbm.so (all errors 0); its device leg runs for the first time in the end-of-round suite."""
import json
import os
import subprocess

import pytest

from _oracle import ref_binary

pytestmark = pytest.mark.gpu

THREADS = max(1, len(os.sched_getaffinity(0)))

RUNS = [
    ("dropin_sc_contact2d", ["nx=96", "ny=48", "steps=1000", "omega=1.0", "RR=14"]),
    ("dropin_hcz_laplace3d", ["nx=10", "ny=8", "nz=12", "steps=40", "omega=1.3", "gravity=-1e-5"]),
    ("dropin_hcz_laplace3d", ["nx=12", "steps=100", "omega=0.5617977528089888"]),     # the shipped config's relaxation rate (ulb .01, Re 6)
    # last: an ODD step count with walls -- the downloaded buffer is then the one the caller did not upload, whose bounce_back
    # nodes must still hold the reference's value-initialised zeros (a combination no earlier Shan-Chen test compares node by node)
    ("dropin_sc_contact2d", ["nx=64", "ny=40", "steps=301", "omega=1.25", "RR=10"]),
]


@pytest.mark.parametrize("exe,args", RUNS)
def test_reference_state_advanced_through_the_c_abi_late(exe, args):
    path = ref_binary(exe)
    if path is None:
        pytest.skip("oracle/_ref/%s not built (needs /root/reference at build time)" % exe)
    r = subprocess.run([path] + args + ["threads=%d" % THREADS], capture_output=True, text=True, timeout=900)
    assert r.returncode in (0, 1), "harness crashed (rc %d): %s" % (r.returncode, r.stderr[-2000:])
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["ok"] and r.returncode == 0, line
