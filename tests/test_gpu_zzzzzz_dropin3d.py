"""Drop-in harness for the HCZ D3Q19 half of north_star's metric: oracle/_ref/dropin_hcz_laplace3d #includes the UNTOUCHED
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
    ["nx=10", "ny=8", "nz=12", "steps=40", "omega=1.3", "gravity=-1e-5"],
    ["nx=12", "steps=100", "omega=0.5617977528089888"],        # the shipped config's relaxation rate (ulb .01, Re 6)
]


@pytest.mark.parametrize("args", RUNS)
def test_reference_3d_state_advanced_through_the_c_abi(args):
    path = ref_binary("dropin_hcz_laplace3d")
    if path is None:
        pytest.skip("oracle/_ref/dropin_hcz_laplace3d not built (needs /root/reference at build time)")
    r = subprocess.run([path] + args + ["threads=%d" % THREADS], capture_output=True, text=True, timeout=900)
    assert r.returncode in (0, 1), "harness crashed (rc %d): %s" % (r.returncode, r.stderr[-2000:])
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["ok"] and r.returncode == 0, line
