"""GPU parity of the compliant-vessel path (CLBM_MODEL_PULSATILE) through the C ABI -- BIT-EXACT.

The device code of this model is built without FMA contraction and keeps the reference's operation order, so the
stored fields P, Ux, Uy, the wall positions, both lattice buffers and the integer node mask must equal the CPU oracle
(oracle/pulsatile_oracle.c, itself pinned bit-for-bit to the reference) and the committed reference dumps exactly.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import _cases
from _oracle import PulsatileOracle, pulsatile_write_vtk

pytestmark = pytest.mark.gpu
clbm = _cases.pkg.clbm


def _compare(dev, ref_fields, ref_lattice=None, ref_parity=None, what=""):
    f = dev.fields()
    np.testing.assert_array_equal(f["flag"], ref_fields["flag"], err_msg=what + " node mask")
    for k in ("yr1", "yr2", "P", "Ux", "Uy"):
        np.testing.assert_array_equal(f[k], ref_fields[k], err_msg="%s %s" % (what, k))
    if ref_lattice is not None:
        lat, par = dev.lattice()
        assert par == ref_parity
        np.testing.assert_array_equal(lat, ref_lattice, err_msg=what + " lattice")


@pytest.mark.parametrize("name", _cases.pulsatile_golden_names())
def test_pulsatile_gpu_vs_reference_dumps(name):
    z, N, dumps, kw = _cases.load_pulsatile_golden(name)
    with clbm.Pulsatile(N=N, **kw) as dev:
        t = 0
        for d in dumps:
            dev.step(d - t)
            t = d
            ref = {k: z["%s_%d" % (k, d)] for k in ("P", "Ux", "Uy", "yr1", "yr2", "flag")}
            lat = z["lattice_%d" % d] if "lattice_%d" % d in z else None
            _compare(dev, ref, lat, int(z["parity_%d" % d]), "%s step %d" % (name, d))


@pytest.mark.parametrize("N,steps,kw", [(64, 1000, {}), (96, 1000, dict(tau=0.9)), (48, 3000, {}), (40, 500, dict(deformable=0))])
def test_pulsatile_gpu_vs_oracle(N, steps, kw):
    """1000+ steps against the oracle on the same initial state, checked along the way"""
    o = PulsatileOracle(N=N, **kw)
    with clbm.Pulsatile(N=N, **kw) as dev:
        assert (dev.nx, dev.ny, dev.tf) == (o.nx, o.ny, o.tf)
        _compare(dev, o.fields(), o.lattice(), o.parity, "initial state")
        done = 0
        for chunk in (1, 1, 8, steps // 2 - 10, steps - steps // 2):
            o.step(chunk)
            dev.step(chunk)
            done += chunk
            _compare(dev, o.fields(), o.lattice(), o.parity, "N=%d step %d" % (N, done))
    o.close()


def test_pulsatile_gpu_reproduces_shipped_vtk(tmp_path):
    """the reference's own golden output (103 VTK files, N = 64) from the device path, byte for byte"""
    hashes = json.load(open(os.path.join(_cases.GOLDEN, "pulsatile_vtk_sha256.json")))
    path = str(tmp_path / "sol.vtk")
    with clbm.Pulsatile(N=64) as dev:
        every = max(1, dev.tf // 100)
        seen = 0
        for t in range(0, dev.tf + 1, every):
            dev.step(t + 1 - dev.t_iter)
            f = dev.fields()
            pulsatile_write_vtk(dev.nx, dev.ny, f["P"], f["Ux"], f["Uy"], f["flag"], t, path)
            assert hashlib.sha256(open(path, "rb").read()).hexdigest() == hashes["sol_%07d.vtk" % t], t
            seen += 1
        assert seen == 103


def test_pulsatile_upload_roundtrip_and_errors():
    o = PulsatileOracle(N=24)
    o.step(37)
    f = o.fields()
    with clbm.Pulsatile(N=24) as dev:
        dev.upload(o.lattice(), f["flag"], f["P"], f["Ux"], f["Uy"], f["yr1"], f["yr2"], o.parity, 37)
        o.step(200)
        dev.step(200)
        _compare(dev, o.fields(), o.lattice(), o.parity, "after upload")
    o.close()
    with pytest.raises(clbm.ClbmError) as e:
        clbm.Pulsatile(N=16, p0_in=0.1, p0_out=0.5, alpha=0.01, is_severed=0)
    assert "Initial wall location out of bounds" in str(e.value)


@pytest.mark.parametrize("N,steps,margin", [(128, 1000, 6.0), (256, 300, 6.0), (96, 1000, 0.0)])
def test_pulsatile_gpu_open_vessel_vs_oracle(N, steps, margin):
    """larger lattices than the reference's own start survives: open vessel at rest, dilating from the
    inlet; 1000 iterations, bit-exact against the oracle, moving walls and fresh nodes throughout"""
    st = _cases.pkg.pulsatile_cases.open_vessel_at_rest(N, margin=margin)
    o = PulsatileOracle(N=N)
    o.set_state(st["lattice"], st["P"], st["Ux"], st["Uy"], st["yr1"], st["yr2"])
    np.testing.assert_array_equal(o.fields()["flag"], st["flag"])
    with clbm.Pulsatile(N=N) as dev:
        dev.upload(st["lattice"], st["flag"], st["P"], st["Ux"], st["Uy"], st["yr1"], st["yr2"], 0, 0)
        done = 0
        for chunk in (1, 9, steps // 2 - 10, steps - steps // 2):
            o.step(chunk)
            dev.step(chunk)
            done += chunk
            _compare(dev, o.fields(), o.lattice(), o.parity, "open vessel N=%d step %d" % (N, done))
        f = dev.fields()
        assert np.isfinite(f["P"]).all() and 0.5 < f["flag"].mean() < 1.0
    o.close()


@pytest.mark.parametrize("variant", [2, 3, 4, 5])
def test_pulsatile_tma_staged_fused_step_is_bit_exact(variant):
    """CLBM_PULS_TMA (opt-in): puls_fused with its 12 input columns staged by cp.async.bulk.tensor boxes; same arithmetic, so the
    state must equal the oracle bit for bit -- moving walls, fresh nodes and both Zou/He ends included (N = 64, the reference's own
    start, 600 iterations, several tiles in y at the 64-row shape)"""
    os.environ["CLBM_PULS_TMA"] = str(variant)
    try:
        o = PulsatileOracle(N=64)
        with clbm.Pulsatile(N=64) as dev:
            for chunk in (1, 2, 297, 300):
                o.step(chunk)
                dev.step(chunk)
                _compare(dev, o.fields(), o.lattice(), o.parity, "variant %d" % variant)
        o.close()
    finally:
        os.environ.pop("CLBM_PULS_TMA", None)
