"""oracle/pulsatile_oracle.c must reproduce the reference's compliant-vessel case bit-for-bit.

(a) binary dumps of the untouched reference header (oracle/_ref/ref_pulsatile -> tests/golden/pulsatile_*.npz, made by
    tests/golden/make_golden_pulsatile.py): P, Ux, Uy, wall positions, node mask, both lattice buffers, parity -- exact.
(b) the reference's OWN golden output: the 103 legacy-VTK files shipped in
    "Abbashub LBM/out_single-phase fluid flow through a compliant vessel/" (N = 64, tf = 2765, one file every 27 steps),
    compared byte-for-byte through their SHA-256 (tests/golden/pulsatile_vtk_sha256.json).
"""
import hashlib
import json
import os

import numpy as np
import pytest

import _cases
from _oracle import PulsatileOracle, pulsatile_write_vtk


@pytest.mark.parametrize("name", _cases.pulsatile_golden_names())
def test_pulsatile_oracle_bit_exact_vs_reference(name):
    z, N, dumps, kw = _cases.load_pulsatile_golden(name)
    o = PulsatileOracle(N=N, **kw)
    t = 0
    for d in dumps:
        o.step(d - t)
        t = d
        f = o.fields()
        np.testing.assert_array_equal(f["flag"], z["flag_%d" % d])     # integer node mask: bit exact
        for k in ("P", "Ux", "Uy", "yr1", "yr2"):
            np.testing.assert_array_equal(f[k], z["%s_%d" % (k, d)], err_msg="%s %s step %d" % (name, k, d))
        assert o.parity == int(z["parity_%d" % d])
        if "lattice_%d" % d in z:
            np.testing.assert_array_equal(o.lattice(), z["lattice_%d" % d])
    o.close()


def test_pulsatile_oracle_reproduces_shipped_vtk(tmp_path):
    hashes = json.load(open(os.path.join(_cases.GOLDEN, "pulsatile_vtk_sha256.json")))
    o = PulsatileOracle(N=64)
    tf = o.tf
    assert tf == 2765
    every = max(1, tf // 100)
    path = str(tmp_path / "sol.vtk")
    seen = 0
    for t in range(tf + 1):
        o.step(1)           # the reference dumps inside iteration t, after the wall update, before the parity flip
        if t % every == 0:
            f = o.fields()
            pulsatile_write_vtk(o.nx, o.ny, f["P"], f["Ux"], f["Uy"], f["flag"], t, path)
            assert hashlib.sha256(open(path, "rb").read()).hexdigest() == hashes["sol_%07d.vtk" % t], t
            seen += 1
    assert seen == len(hashes) == 103
    o.close()


def test_pulsatile_initial_wall_out_of_bounds():
    """AB/apps/PulsatileBloodFlow2D.h:181 throws runtime_error("Initial wall location out of bounds.")"""
    with pytest.raises(RuntimeError):
        PulsatileOracle(N=16, p0_in=0.1, p0_out=0.5, alpha=0.01, is_severed=0)


@pytest.mark.parametrize("name", ["open_N128_m6", "open_N256_m6"])
def test_pulsatile_oracle_bit_exact_vs_reference_from_the_open_vessel(name):
    """the start state of the large-N device tests and of bench.py's N = 1024 line (pulsatile_cases.open_vessel_at_rest) handed
    to the UNTOUCHED reference header at N = 128 / 256 (tests/golden/make_golden_pulsatile_open.py): the oracle must make of it
    exactly what the reference does -- walls dilating from the inlet, fresh nodes, both Zou/He ends -- SHA-256 of every array."""
    rec = json.load(open(os.path.join(_cases.GOLDEN, "pulsatile_open_sha256.json")))[name]
    N = rec["N"]
    st = _cases.pkg.pulsatile_cases.open_vessel_at_rest(N, margin=rec["margin"])
    o = PulsatileOracle(N=N)
    o.set_state(st["lattice"], st["P"], st["Ux"], st["Uy"], st["yr1"], st["yr2"], 0, 0)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert int((o.fields()["flag"] == 1).sum()) == rec["initial_bulk_nodes"]
    t = 0
    for d in rec["dumps"]:
        o.step(d - t)
        t = d
        want = rec["steps"][str(d)]
        f = o.fields()
        assert int((f["flag"] == 1).sum()) == want["bulk_nodes"]
        assert sha(f["flag"]) == want["flag"]
        for k in ("yr1", "yr2", "P", "Ux", "Uy"):
            assert sha(f[k]) == want[k], (name, d, k)
        assert sha(o.lattice()) == want["lattice"], (name, d)
        assert o.parity == want["parity"]
    last = rec["steps"][str(rec["dumps"][-1])]
    assert last["bulk_nodes"] > rec["initial_bulk_nodes"] and last["yr2_min_max"][1] - last["yr2_min_max"][0] > 5.0   # the walls moved
    o.close()
