"""The drop-in claim, tested where it matters: binaries that #include the UNTOUCHED reference case headers
(oracle/_ref/dropin_*, built by `make -C oracle ref` from tests/host_check/dropin/*.cpp where /root/reference exists, shipped
to the GPU box prebuilt).  The REFERENCE's iniLattice + inigeom_* build the host state; clbm_create / clbm_upload / clbm_step /
clbm_download_lattice advance a copy of it; the REFERENCE's density / u_actual / macro_phi_P / velocity / totalMass / ... are
then evaluated on both arrays and must agree to 1e-10 after 1000 steps (masks, parity and the integer scans exactly).
The code between the "(B)" markers of each harness is the patch INTEGRATION.md shows."""
import json
import os
import subprocess

import pytest

import _cases
from _oracle import ref_binary

pytestmark = pytest.mark.gpu

THREADS = max(1, len(os.sched_getaffinity(0)))

RUNS = [
    ("dropin_sc_laplace2d", ["nx=64", "ny=64", "steps=1000", "omega=1.0"]),
    ("dropin_sc_laplace2d", ["nx=48", "ny=40", "steps=301", "omega=0.5618", "gravity=-1e-5"]),
    ("dropin_hcz_rt2d", ["nx=24", "ny=98", "steps=1000", "omega=1.0"]),
    ("dropin_hcz_rt2d", ["nx=32", "ny=130", "steps=200", "omega=1.7"]),
    ("dropin_hcz_layered2d", ["nx=10", "ny=41", "steps=301"]),
    ("dropin_hcz_layered2d", ["nx=10", "ny=101", "steps=1000", "w_int=4"]),
]


@pytest.mark.parametrize("exe,args", RUNS)
def test_reference_state_advanced_through_the_c_abi(exe, args):
    path = ref_binary(exe)
    if path is None:
        pytest.skip("oracle/_ref/%s not built (needs /root/reference at build time)" % exe)
    r = subprocess.run([path] + args + ["threads=%d" % THREADS], capture_output=True, text=True, timeout=900)
    assert r.returncode in (0, 1), "harness crashed (rc %d): %s" % (r.returncode, r.stderr[-2000:])
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["ok"] and r.returncode == 0, line
