"""Boundary-first overlap protocol (clbm_step_stage 10-12) for the HCZ D2Q9 fused kernel: the interior columns [2, nx-2) on the
launching stream, the 2 + 2 boundary columns (phi stencil chain of reach 2) and both exchanges on the boundary stream.  Must be
bit-identical to the sequential protocol (stages 0-2) and to the single slab -- BASELINE configs[2] (2048 x 8194 over 8 GPUs)
is the case it exists for.  One GPU, R slab contexts, device-to-device copies as the exchange (LocalRing)."""
import numpy as np
import pytest

import _cases

pytestmark = pytest.mark.gpu
pkg = _cases.pkg
P = pkg.params
slab = pkg.slab


def _params(kind, nx, ny):
    if kind == "mrt":
        return P.hcz_mrt_params(nx, ny, N=256, s_e=1.8, s_eps=1.9, s_q=1.7)
    if kind == "layered":
        return P.hcz_layered_params(nx, ny, gx_const=1e-6)
    return P.hcz_params(P.MODEL_HCZ_D2Q9, nx, ny, N=256)


@pytest.mark.parametrize("kind,nx,ny,nranks", [("bgk", 32, 130, 2), ("bgk", 32, 130, 4), ("bgk", 16, 300, 4),     # 4 columns per slab: no interior
                                               ("bgk", 21, 66, 3), ("mrt", 24, 130, 2), ("layered", 20, 41, 2)])
def test_hcz2d_overlap_protocol_is_bit_identical(kind, nx, ny, nranks):
    prm = _params(kind, nx, ny)
    case, args = (P.CASE_HCZ_LAYERED2D, (0.3, 2.0)) if kind == "layered" else (P.CASE_HCZ_RT2D, ())
    steps = 40
    with pkg.clbm.Lattice(prm) as single:
        single.init_case(case, args)
        single.step(steps)
        ref = single.in_pops()
    out = {}
    for overlap in (False, True):
        lats = [pkg.clbm.Lattice(slab.slab_params(prm, r, nranks)) for r in range(nranks)]
        assert all(lat.overlap_supported() for lat in lats)
        for lat in lats:
            lat.init_case(case, args)
        ring = slab.LocalRing(lats)
        ring.step(steps // 2, overlap=overlap)
        ring.step(steps - steps // 2, overlap=not overlap)      # the two protocols can be mixed between steps
        out[overlap] = np.concatenate([lat.in_pops() for lat in lats], axis=2)
        for lat in lats:
            lat.close()
    np.testing.assert_array_equal(out[False], ref)
    np.testing.assert_array_equal(out[True], ref)


def test_hcz2d_overlap_needs_the_fused_kernel_and_four_columns():
    prm = P.hcz_params(P.MODEL_HCZ_D2Q9, 32, 66, N=32)
    with pkg.clbm.Lattice(slab.slab_params(prm.copy(fused=0), 0, 2)) as lat:
        assert not lat.overlap_supported()                      # staged kernels: sequential protocol only
    with pkg.clbm.Lattice(slab.slab_params(P.hcz_params(P.MODEL_HCZ_D2Q9, 9, 66, N=32), 0, 3)) as lat:
        assert not lat.overlap_supported()                      # 3 columns per slab
    with pkg.clbm.Lattice(slab.slab_params(prm, 1, 2)) as lat:
        assert lat.overlap_supported()
