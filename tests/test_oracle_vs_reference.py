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
