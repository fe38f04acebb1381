"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/clbm.h declares,
rejects bad arguments, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import _cases

pkg = _cases.pkg
P = pkg.params
ROOT = _cases.ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "clbm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(clbm_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_all_exported():
    lib = pkg.clbm.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "libclbm.so does not export %s" % s
    assert sorted(pkg.clbm.EXPORTS) == syms
    assert lib.clbm_abi_version() == P.ABI_VERSION


def test_params_struct_matches_header_layout():
    # 10 int32 + 17 doubles (ABI 2) + 3 doubles + 2 int32 (appended in ABI 3), naturally aligned: sizeof(clbm_params) on the C side
    assert ctypes.sizeof(P.Params) == 10 * 4 + 17 * 8 + 3 * 8 + 2 * 4
    assert P.Params.s_e.offset == 10 * 4 + 17 * 8 and P.Params.collision.offset == 10 * 4 + 20 * 8
    assert P.Params.omega.offset == 40


def test_bad_arguments_are_rejected_without_a_device_call():
    lib = pkg.clbm.load_library()
    h = ctypes.c_void_p()
    bad = P.make_params(P.MODEL_SC_D2Q9, 8, 8)
    bad.abi_version = 99
    assert lib.clbm_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    assert b"ABI" in lib.clbm_last_error()
    bad = P.make_params(P.MODEL_SC_D2Q9, 8, 8, nz=4)
    assert lib.clbm_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    bad = P.make_params(7, 8, 8)
    assert lib.clbm_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    assert lib.clbm_step(None, 1) == -1
    assert lib.clbm_destroy(None) == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.clbm.ClbmError) as e:
        pkg.clbm.Lattice(P.sc_params(P.MODEL_SC_D2Q9, 8, 8))
    assert "no CUDA device" in str(e.value)


def test_missing_extension_fails_loudly(tmp_path):
    with pytest.raises(pkg.clbm.ClbmError):
        pkg.clbm.load_library(str(tmp_path / "libclbm.so"))


def test_library_driven_ring_entry_points_fail_cleanly():
    """clbm_comm_* / clbm_slab_step (csrc/slab_comm.cu): NCCL is resolved at run time, bad arguments are rejected before any
    device or NCCL call, nothing crashes in a container without a GPU"""
    lib = pkg.clbm.load_library()
    buf = ctypes.create_string_buffer(128)
    rc = lib.clbm_comm_unique_id(buf)
    assert rc in (0, -5, -3)                 # ok, or "NCCL not found" / an NCCL error, reported through clbm_last_error
    if rc == 0:
        assert any(b != 0 for b in buf.raw)
    else:
        assert lib.clbm_last_error()
    assert lib.clbm_comm_unique_id(None) == -1
    assert lib.clbm_slab_step(None, 1) == -1
    assert lib.clbm_comm_init(None, buf, 0, 2) == -1
    assert lib.clbm_comm_destroy(None) == 0


def test_every_entry_point_survives_null_arguments():
    """nothing crosses the ABI as a crash: every export called with a null context / null pointers / zeros returns an error code
    (or a benign 0 / NULL for the destroy, query and stream getters) -- on a box without a device as well"""
    lib = pkg.clbm.load_library()
    benign_zero = {"clbm_destroy", "clbm_comm_destroy", "clbm_peer_disconnect", "clbm_pulsatile_destroy", "clbm_yl2d_destroy",
                   "clbm_overlap_supported", "clbm_overlap_variant", "clbm_overlap_width", "clbm_ring_kind",
                   "clbm_pulsatile_launch_count", "clbm_yl2d_launch_count"}
    not_swept = {"clbm_last_error", "clbm_abi_version", "clbm_alloc_host", "clbm_free_host", "clbm_comm_unique_id"}
    swept = 0
    for name in sorted(pkg.clbm.EXPORTS):
        if name in not_swept:
            continue
        fn = getattr(lib, name)
        assert fn.argtypes is not None, name + " has no ctypes signature"
        args = [0 if t in (ctypes.c_int, ctypes.c_int64, ctypes.c_size_t) else 0.0 if t in (ctypes.c_double, ctypes.c_float) else None
                for t in fn.argtypes]
        rc = fn(*args)
        swept += 1
        if name in ("clbm_stream", "clbm_boundary_stream"):
            assert rc is None, name
        elif name in benign_zero:
            assert rc == 0, (name, rc)
        else:
            assert rc < 0, (name, rc)
            assert lib.clbm_last_error(), name
    assert swept >= 55
    assert lib.clbm_free_host(None) == 0 and lib.clbm_alloc_host(64, None) == -1      # free(NULL) is a no-op, as in C


def test_every_environment_knob_is_documented():
    """INTEGRATION.md's tuning-knob table lists every CLBM_* / COOLBM_* variable the sources read"""
    import glob
    names = set()
    pkgdir = os.path.join(ROOT, "multiphase-lbm_b200")
    for pat in ("csrc/*.cu", "csrc/*.cuh", "csrc/*.h", "apps/*.h", "apps/*.cpp", "*.py"):
        for f in glob.glob(os.path.join(pkgdir, pat)):
            names |= set(re.findall(r'"((?:CLBM|COOLBM)_[A-Z0-9_]+)"', open(f, errors="replace").read()))
    assert len(names) >= 25
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read() + open(os.path.join(ROOT, "README.md")).read()
    missing = sorted(n for n in names if n not in doc)
    assert not missing, missing
