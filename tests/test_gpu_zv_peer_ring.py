"""The peer-memory ring (csrc/slab_comm.cu) on ONE GPU.  Contexts of this process are connected with clbm_peer_connect_local,
so every pack writes straight into the neighbour's mailbox and the exchanges are the flag signal / wait kernels -- the code
path of a multi-GPU run minus the IPC mapping (tools/slab_check.py and bench.py's slab_bit_identical cover that under
torchrun).  Results must equal the single slab bit for bit."""
import os

import numpy as np
import pytest

import _cases

pytestmark = pytest.mark.gpu
pkg = _cases.pkg
P = pkg.params
slab = pkg.slab

CASES = {
    "sc2d_contact": (lambda: P.sc_params(P.MODEL_SC_D2Q9, 48, 32, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT),
                     P.CASE_SC_CONTACT2D, (0.265, 0.038, 9.0), 61),
    "sc3d_sessile": (lambda: P.sc_params(P.MODEL_SC_D3Q19, 24, 16, 20, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT),
                     P.CASE_SC_DROPLET3D, (0.265, 0.038, 6.0, 5.0), 41),
    "hcz2d_rt": (lambda: P.hcz_params(P.MODEL_HCZ_D2Q9, 32, 66, N=32), P.CASE_HCZ_RT2D, (), 41),
    "hcz3d_drop": (lambda: P.hcz_params(P.MODEL_HCZ_D3Q19, 24, 12, 12, ulb=0.01, N=24, Re=6.0, kappa=5e-4, gravity=-1e-5),
                   P.CASE_HCZ_LAPLACE3D, (), 25),
}


def single_run(prm, case, args, steps):
    with pkg.clbm.Lattice(prm) as single:
        single.init_case(case, args)
        single.step(steps)
        return single.in_pops(), single.fields()


def make_ring(prm, case, args, nranks):
    lats = [pkg.clbm.Lattice(slab.slab_params(prm, r, nranks)) for r in range(nranks)]
    for lat in lats:
        lat.init_case(case, args)
    ring = slab.LocalRing(lats, peer=True)
    assert all(lat.ring_kind() == 2 for lat in lats)
    return lats, ring


@pytest.mark.parametrize("nranks", [2, 3])
@pytest.mark.parametrize("name", sorted(CASES))
def test_peer_ring_matches_single_slab(name, nranks):
    """R contexts of this process on a peer ring: packs into the neighbours' mailboxes, flag signal / wait kernels, no host
    synchronisation (the stages of the contexts are interleaved by LocalRing: all signals of a phase before any wait)"""
    mk, case, args, steps = CASES[name]
    prm = mk()
    ref_pops, ref_fields = single_run(prm, case, args, steps)
    lats, ring = make_ring(prm, case, args, nranks)
    ring.step(steps)
    ring.refresh_moment_halo()
    pops = np.concatenate([lat.in_pops() for lat in lats], axis=2)
    fields = {k: np.concatenate([lat.fields()[k] for lat in lats]) for k in ("s0", "ux", "uy", "uz")}
    for lat in lats:
        lat.close()
    np.testing.assert_array_equal(pops, ref_pops)
    for k in fields:
        np.testing.assert_array_equal(fields[k], ref_fields[k])


class env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update({k: str(v) for k, v in self.kw.items()})

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("overlap", [-1, 0, 1, 2])
@pytest.mark.parametrize("graph", [1, 0])
@pytest.mark.parametrize("name", sorted(CASES))
def test_self_ring_graph_replay_matches_single_slab(name, graph, overlap):
    """clbm_slab_step itself -- whole steps issued by the library, two steps per CUDA-graph launch from the third step on --
    on a ring of ONE context that is its own neighbour (CLBM_FORCE_SLAB=1: the full-width lattice runs in x-slab mode, its
    ghost planes filled through its own mailbox).  This is the code path of a one-GPU-per-process run, without a second
    context on the GPU whose signal a spinning wait could starve.  Bit-identical to the plain single-slab run."""
    mk, case, args, steps = CASES[name]
    prm = mk()
    ref_pops, ref_fields = single_run(prm, case, args, steps)
    if overlap > 0 and name in ("sc2d_contact", "hcz3d_drop"):
        pytest.skip("no overlap protocol for this kernel")
    kw = dict(CLBM_FORCE_SLAB=1, CLBM_SLAB_GRAPH=graph)       # read once, in clbm_create
    if overlap >= 0:
        kw["CLBM_SLAB_OVERLAP"] = overlap                        # -1: the model's default protocol
    with env(**kw):
        lat = pkg.clbm.Lattice(prm)
    lat.init_case(case, args)
    lat.peer_connect_local(lat, lat)
    assert lat.ring_kind() == 2
    if overlap >= 0:
        assert lat.overlap_variant() == overlap
    l0 = lat.launch_count()
    for n in (1, 7, 2, 9, steps - 19):          # eager first steps, graph replays, odd leftovers, both parities
        lat.slab_step(n)
    per_step = (lat.launch_count() - l0) / steps
    lat.step_stage(20)
    lat.slab_exchange(0)
    pops, fields = lat.in_pops(), lat.fields()
    lat.close()
    np.testing.assert_array_equal(pops, ref_pops)
    for k in ("s0", "ux", "uy", "uz"):
        np.testing.assert_array_equal(fields[k], ref_fields[k])
    assert 3 <= per_step <= 16, per_step       # replayed launches are counted too


@pytest.mark.parametrize("form", [1, 2])
def test_protocols_mix_without_host_synchronisation(form):
    """steps of the sequential protocol (stages 0-2 on the launching stream) between steps of the overlap protocol (stages
    10-12, boundary stream), every call asynchronous: the cross-stream ordering must come from events, not from a host sync
    (the boundary stream has to wait for what the launching stream ran last)"""
    prm = P.sc_params(P.MODEL_SC_D3Q19, 32, 24, 36, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT)
    case, args = P.CASE_SC_DROPLET3D, (0.265, 0.038, 8.0, 5.0)
    ref_pops, _ = single_run(prm, case, args, 30)
    with env(CLBM_SLAB_OVERLAP=form):            # the Shan-Chen default is the sequential protocol
        lats, ring = make_ring(prm, case, args, 2)
    assert all(lat.overlap_supported() and lat.overlap_variant() == form for lat in lats)
    for _ in range(5):
        ring.step(3, overlap=True)
        ring.step(1, overlap=False)
        ring.step(1, overlap=True)
        ring.step(1, overlap=False)
    pops = np.concatenate([lat.in_pops() for lat in lats], axis=2)
    for lat in lats:
        lat.close()
    np.testing.assert_array_equal(pops, ref_pops)


def test_peer_ring_full_plane_slabs_hcz2d_config3_shape():
    """BASELINE configs[2] plane shape: 4 slabs of 16 columns x 8194 rows, overlap protocol, against the single 64 x 8194 slab,
    bit for bit"""
    prm = P.hcz_params(P.MODEL_HCZ_D2Q9, 64, 8194, 1, ulb=0.04, N=2048, Re=3000.0)
    ref_pops, _ = single_run(prm, P.CASE_HCZ_RT2D, (), 12)
    lats, ring = make_ring(prm, P.CASE_HCZ_RT2D, (), 4)
    ring.step(12)
    pops = np.concatenate([lat.in_pops() for lat in lats], axis=2)
    for lat in lats:
        lat.close()
    np.testing.assert_array_equal(pops, ref_pops)
