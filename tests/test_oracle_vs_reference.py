"""The oracle restatement must reproduce the UNTOUCHED reference functor bit-for-bit.

tests/golden/*.npz were produced by oracle/_ref (the reference headers compiled where they
lie, see tests/golden/make_golden.py).  Populations and macroscopic fields are compared with
exact equality: the restatement keeps the reference's expressions and summation order and is
built with -ffp-contract=off.
"""
import numpy as np
import pytest

import _cases
from _oracle import OracleSim

NAMES = _cases.golden_names(long_horizon=True)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_bit_exact_vs_reference(name):
    z, _ = _cases.load_golden(name)
    p, case_id, args, steps, fmap = _cases.golden_setup(name)
    sim = OracleSim(p).init_case(case_id, args)
    np.testing.assert_array_equal(sim.flag, z["flag"])          # integer masks: bit exact
    sim.step(steps, threads=1)
    pops = sim.in_pops()
    assert pops.shape == z["pops"].shape
    np.testing.assert_array_equal(pops, z["pops"])
    f = sim.fields()
    for gname, slot in fmap.items():
        np.testing.assert_array_equal(f[slot], z[gname], err_msg="%s field %s" % (name, gname))
    if "fx" in z.files:     # fixtures that also hold the reference's force_ff (SC/apps/RayleighTaylor2D.h:236-289)
        F = sim.force()
        np.testing.assert_array_equal(F["fx"], z["fx"])
        np.testing.assert_array_equal(F["fy"], z["fy"])


@pytest.mark.parametrize("name", NAMES[:1] + NAMES[-1:])
def test_oracle_thread_count_invariant(name):
    """OpenMP sharding must not change a single bit (race-free two-lattice push)."""
    p, case_id, args, steps, _ = _cases.golden_setup(name)
    a = OracleSim(p).init_case(case_id, args).step(min(steps, 10), threads=1)
    b = OracleSim(p).init_case(case_id, args).step(min(steps, 10), threads=4)
    np.testing.assert_array_equal(a.lattice, b.lattice)


@pytest.mark.parametrize("name", ["c1_sc_d2q9_256", "c2_hcz_d2q9_256", "c3_shape_16x8194"])
def test_oracle_bit_exact_vs_reference_on_full_baseline_configs(name):
    """BASELINE.json configs[0] (Shan-Chen 256 x 256) and configs[1] (HCZ Rayleigh-Taylor 256 x 1026) in full -- size, shipped
    parameters, 1000 steps -- and the column shape of configs[2] (8194 rows, N = 2048 parameters, 16 columns, 100 steps; the
    whole 2048 x 8194 lattice is 36 h of the functor) through the UNTOUCHED functors (tests/golden/make_golden_baseline_configs.py; SHA-256 of the
    populations, every field and the mask, the arrays being tens of MB).  The GPU suite holds the device to the oracle on exactly
    these two configurations (test_sc_laplace2d_256_1000_steps, the 256 x 1026 1000-step test)."""
    import hashlib
    import json
    import os
    rec = json.load(open(os.path.join(_cases.GOLDEN, "baseline_configs_sha256.json")))[name]
    kw = rec["params"]
    P = _cases.P
    if name.startswith("c1"):
        p = P.sc_params(P.MODEL_SC_D2Q9, 256, 256, ulb=0.01, N=256, Re=6.0)            # as test_gpu_parity.py builds configs[0]
        case_id, args, fmap = P.CASE_SC_LAPLACE2D, (kw["rhol"], kw["rhog"], 10.0), {"rho": "s0", "pressure": "s1", "ux": "ux", "uy": "uy"}
    else:
        p = P.hcz_params(P.MODEL_HCZ_D2Q9, 256, 1026, N=256) if name.startswith("c2") else \
            P.hcz_params(P.MODEL_HCZ_D2Q9, 16, 8194, 1, ulb=0.04, N=2048, Re=3000.0)    # as test_gpu_parity.py builds configs[1] / the configs[2] column shape
        case_id, args, fmap = P.CASE_HCZ_RT2D, (), {"phi": "s0", "P": "s1", "rho": "s2", "ux": "ux", "uy": "uy"}
    assert (p.nx, p.ny, p.omega) == (kw["nx"], kw["ny"], kw["omega"])
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    sim = OracleSim(p).init_case(case_id, args)
    assert sha(sim.flag) == rec["sha256"]["flag"] and int((sim.flag == 1).sum()) == rec["bulk_nodes"]
    sim.step(kw["steps"])                                  # all host threads: the thread count does not change a bit (test above)
    assert sha(sim.in_pops()) == rec["sha256"]["pops"]
    f = sim.fields()
    for gname, slot in fmap.items():
        assert float(np.abs(f[slot]).max()) == rec["max_abs"][gname], gname
        assert sha(f[slot]) == rec["sha256"][gname], gname
