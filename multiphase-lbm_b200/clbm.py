"""ctypes binding of the clbm C ABI (include/clbm.h) -- the host-side mirror of the reference's
LBM_* functor object: raw lattice / flag / parity on the host side, the time step on the device.

There is no CPU fallback: if csrc/libclbm.so is missing or no CUDA device is present, creating a
Lattice raises.  Build the library with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C multiphase-lbm_b200/csrc`.
"""
import ctypes
import os

import numpy as np

from . import params as P

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libclbm.so")

EXPORTS = [
    "clbm_create", "clbm_destroy", "clbm_last_error", "clbm_abi_version", "clbm_upload", "clbm_upload2", "clbm_download_lattice",
    "clbm_download_fields", "clbm_download_force", "clbm_init_case", "clbm_step", "clbm_sync", "clbm_step_timed", "clbm_launch_count",
    "clbm_profile_step", "clbm_reduce", "clbm_diag_contact_angle", "clbm_diag_contact_angle_slab", "clbm_diag_interface_heights", "clbm_halo_buffer", "clbm_halo_pack", "clbm_halo_unpack", "clbm_step_stage",
    "clbm_stream", "clbm_overlap_supported", "clbm_overlap_variant", "clbm_overlap_width", "clbm_boundary_stream", "clbm_comm_unique_id", "clbm_comm_init", "clbm_slab_step", "clbm_comm_destroy",
    "clbm_peer_export", "clbm_peer_connect", "clbm_peer_connect_local", "clbm_peer_disconnect", "clbm_ring_kind", "clbm_slab_exchange", "clbm_slab_signal", "clbm_slab_wait", "clbm_kernel_timing_begin", "clbm_kernel_timing_end", "clbm_alloc_host", "clbm_free_host",
    "clbm_pulsatile_create", "clbm_pulsatile_destroy", "clbm_pulsatile_info", "clbm_pulsatile_step",
    "clbm_pulsatile_step_timed", "clbm_pulsatile_sync", "clbm_pulsatile_launch_count",
    "clbm_pulsatile_kernel_timing_begin", "clbm_pulsatile_kernel_timing_end", "clbm_pulsatile_download_fields",
    "clbm_pulsatile_download_lattice", "clbm_pulsatile_upload",
    "clbm_yl2d_create", "clbm_yl2d_destroy", "clbm_yl2d_step", "clbm_yl2d_step_timed", "clbm_yl2d_sync",
    "clbm_yl2d_launch_count", "clbm_yl2d_download_fields", "clbm_yl2d_download_lattice", "clbm_yl2d_upload", "clbm_yl2d_reduce",
]

_lib = None


class ClbmError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen libclbm.so and declare the prototypes.  Fails loudly when the extension is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise ClbmError("CUDA extension %s is missing: build it first (there is no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    vp, dp, u8p = ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint8)
    lib.clbm_create.argtypes = [ctypes.POINTER(P.Params), ctypes.POINTER(vp)]
    lib.clbm_destroy.argtypes = [vp]
    lib.clbm_last_error.restype = ctypes.c_char_p
    lib.clbm_upload.argtypes = [vp, vp, vp, ctypes.c_int]
    lib.clbm_upload2.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_int]
    lib.clbm_download_lattice.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_int)]
    lib.clbm_download_fields.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    lib.clbm_download_force.argtypes = [vp, vp, vp, vp]
    lib.clbm_init_case.argtypes = [vp, ctypes.c_int, dp, ctypes.c_int]
    lib.clbm_step.argtypes = [vp, ctypes.c_int]
    lib.clbm_sync.argtypes = [vp]
    lib.clbm_step_timed.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
    lib.clbm_launch_count.argtypes = [vp]
    lib.clbm_launch_count.restype = ctypes.c_int64
    lib.clbm_profile_step.argtypes = [vp, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_float), ctypes.c_int]
    lib.clbm_reduce.argtypes = [vp, ctypes.c_int, dp]
    lib.clbm_diag_contact_angle.argtypes = [vp, ctypes.c_double] + [ctypes.POINTER(ctypes.c_int)] * 3
    lib.clbm_diag_contact_angle_slab.argtypes = [vp, ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    lib.clbm_diag_interface_heights.argtypes = [vp, ctypes.c_double] + [ctypes.POINTER(ctypes.c_int)] * 2
    lib.clbm_halo_buffer.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp),
                                     ctypes.POINTER(ctypes.c_size_t)]
    lib.clbm_halo_pack.argtypes = [vp, ctypes.c_int]
    lib.clbm_halo_unpack.argtypes = [vp, ctypes.c_int]
    lib.clbm_step_stage.argtypes = [vp, ctypes.c_int]
    lib.clbm_kernel_timing_begin.argtypes = [vp, ctypes.c_int]
    lib.clbm_kernel_timing_end.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int),
                                           ctypes.POINTER(ctypes.c_char_p)]
    lib.clbm_alloc_host.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.clbm_free_host.argtypes = [vp]
    lib.clbm_stream.argtypes = [vp]
    lib.clbm_stream.restype = vp
    lib.clbm_boundary_stream.argtypes = [vp]
    lib.clbm_boundary_stream.restype = vp
    lib.clbm_overlap_supported.argtypes = [vp]
    lib.clbm_overlap_width.argtypes = [vp]
    lib.clbm_overlap_variant.argtypes = [vp]
    lib.clbm_comm_unique_id.argtypes = [vp]
    lib.clbm_comm_init.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int]
    lib.clbm_slab_step.argtypes = [vp, ctypes.c_int]
    lib.clbm_comm_destroy.argtypes = [vp]
    lib.clbm_peer_export.argtypes = [vp, vp]
    lib.clbm_peer_connect.argtypes = [vp, vp, vp]
    lib.clbm_peer_connect_local.argtypes = [vp, vp, vp]
    lib.clbm_peer_disconnect.argtypes = [vp]
    lib.clbm_ring_kind.argtypes = [vp]
    lib.clbm_slab_exchange.argtypes = [vp, ctypes.c_int]
    lib.clbm_slab_signal.argtypes = [vp, ctypes.c_int, ctypes.c_int]
    lib.clbm_slab_wait.argtypes = [vp, ctypes.c_int, ctypes.c_int]
    ip, fp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_float)
    lib.clbm_pulsatile_create.argtypes = [ctypes.POINTER(P.PulsatileParams), ctypes.POINTER(vp)]
    lib.clbm_pulsatile_destroy.argtypes = [vp]
    lib.clbm_pulsatile_info.argtypes = [vp, ip, ip, ip, ip, ip]
    lib.clbm_pulsatile_step.argtypes = [vp, ctypes.c_int]
    lib.clbm_pulsatile_step_timed.argtypes = [vp, ctypes.c_int, fp]
    lib.clbm_pulsatile_sync.argtypes = [vp]
    lib.clbm_pulsatile_launch_count.argtypes = [vp]
    lib.clbm_pulsatile_launch_count.restype = ctypes.c_int64
    lib.clbm_pulsatile_kernel_timing_begin.argtypes = [vp, ctypes.c_int]
    lib.clbm_pulsatile_kernel_timing_end.argtypes = [vp, fp, ip]
    lib.clbm_pulsatile_download_fields.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.clbm_pulsatile_download_lattice.argtypes = [vp, vp, ip]
    lib.clbm_pulsatile_upload.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int]
    lib.clbm_yl2d_create.argtypes = [ctypes.POINTER(P.YL2DParams), ctypes.POINTER(vp)]
    lib.clbm_yl2d_destroy.argtypes = [vp]
    lib.clbm_yl2d_step.argtypes = [vp, ctypes.c_int]
    lib.clbm_yl2d_step_timed.argtypes = [vp, ctypes.c_int, fp]
    lib.clbm_yl2d_sync.argtypes = [vp]
    lib.clbm_yl2d_launch_count.argtypes = [vp]
    lib.clbm_yl2d_launch_count.restype = ctypes.c_int64
    lib.clbm_yl2d_download_fields.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.clbm_yl2d_download_lattice.argtypes = [vp, vp, ip]
    lib.clbm_yl2d_upload.argtypes = [vp, vp, vp, vp, ctypes.c_int]
    lib.clbm_yl2d_reduce.argtypes = [vp, ctypes.c_int, dp]
    for name in EXPORTS:
        getattr(lib, name)  # every declared symbol must resolve
    _lib = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if hasattr(a, "data_ptr"):   # torch tensor (pinned host memory)
        return a.data_ptr()
    return a


class Lattice:
    """Device-resident lattice of one x-slab; the counterpart of an `LBM_*` functor plus its arrays."""

    def __init__(self, params):
        self.lib = load_library()
        self.p = params
        h = ctypes.c_void_p()
        self._h = None
        self._check(self.lib.clbm_create(ctypes.byref(params), ctypes.byref(h)))
        self._h = h

    # -- plumbing
    def _check(self, rc):
        if rc != 0:
            raise ClbmError("clbm error %d: %s" % (rc, self.lib.clbm_last_error().decode()))

    def close(self):
        if self._h is not None:
            self.lib.clbm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- state transfer (reference layout)
    def upload(self, lattice, flag, parity=0, other_buffer=True):
        """other_buffer=False: `lattice` holds valid data only in the buffer `parity` selects (clbm_upload2)"""
        assert lattice.dtype == np.float64 if isinstance(lattice, np.ndarray) else True
        self._check(self.lib.clbm_upload2(self._h, _ptr(lattice), _ptr(flag), int(parity), 1 if other_buffer else 0))

    def download_lattice(self, lattice=None):
        if lattice is None:
            lattice = np.zeros(self.p.lattice_size, dtype=np.float64)
        par = ctypes.c_int(0)
        self._check(self.lib.clbm_download_lattice(self._h, _ptr(lattice), ctypes.byref(par)))
        return lattice, par.value

    def in_pops(self):
        """current populations as [sets, Q, nelem] (same view as tests/_oracle.OracleSim.in_pops)"""
        lat, par = self.download_lattice()
        p = self.p
        npop = p.Q * p.nelem
        return np.stack([lat[s * 2 * npop + par * npop: s * 2 * npop + (par + 1) * npop].reshape(p.Q, p.nelem)
                         for s in range(p.sets)])

    def fields(self, names=("s0", "s1", "s2", "ux", "uy", "uz"), out=None):
        """macroscopic fields with the reference's definitions"""
        n = self.p.nelem
        order = ["s0", "s1", "s2", "ux", "uy", "uz"]
        arrs = out or {k: np.zeros(n) for k in names}
        ptrs = [_ptr(arrs[k]) if k in arrs else None for k in order]
        self._check(self.lib.clbm_download_fields(self._h, *ptrs, None))
        return arrs

    def force(self):
        """Shan-Chen interaction force of every node (the VECTORS Force of the reference's VTK files)"""
        n = self.p.nelem
        f = [np.zeros(n) for _ in range(3)]
        self._check(self.lib.clbm_download_force(self._h, *[_ptr(a) for a in f]))
        return f

    def flags(self):
        f = np.zeros(self.p.nelem, dtype=np.uint8)
        self._check(self.lib.clbm_download_fields(self._h, None, None, None, None, None, None, _ptr(f)))
        return f

    def init_case(self, case_id, args=()):
        a = np.asarray(args, dtype=np.float64)
        self._check(self.lib.clbm_init_case(self._h, int(case_id), a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                            int(a.size)))
        return self

    # -- the hot path
    def step(self, n=1):
        self._check(self.lib.clbm_step(self._h, int(n)))
        return self

    def sync(self):
        self._check(self.lib.clbm_sync(self._h))

    def step_timed(self, n):
        ms = ctypes.c_float(0)
        self._check(self.lib.clbm_step_timed(self._h, int(n), ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.lib.clbm_launch_count(self._h))

    def profile_step(self):
        cap = 64
        names = (ctypes.c_char_p * cap)()
        ms = (ctypes.c_float * cap)()
        n = self.lib.clbm_profile_step(self._h, names, ms, cap)
        if n < 0:
            self._check(n)
        return [(names[i].decode(), ms[i]) for i in range(n)]

    def kernel_timing_begin(self, cap=256):
        self._check(self.lib.clbm_kernel_timing_begin(self._h, int(cap)))

    def kernel_timing_end(self):
        ms, n, name = ctypes.c_float(0), ctypes.c_int(0), ctypes.c_char_p()
        self._check(self.lib.clbm_kernel_timing_end(self._h, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(name)))
        return ms.value, n.value, (name.value or b"").decode()

    def reduce(self, kind):
        out = ctypes.c_double(0)
        self._check(self.lib.clbm_reduce(self._h, int(kind), ctypes.byref(out)))
        return out.value

    def contact_angle_scan(self, rho_cut):
        """device-side scans of calculateContactAngle (SC/apps/contactAngle2D.h:465-529) -> (base_y, base, height)"""
        v = [ctypes.c_int(0) for _ in range(3)]
        self._check(self.lib.clbm_diag_contact_angle(self._h, ctypes.c_double(rho_cut), *[ctypes.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def contact_angle_scan_slab(self, rho_cut, base_y_in=-1):
        """the partial scan of one x-slab in global x (clbm_diag_contact_angle_slab) -> [base_y, lstop, rstop, hstop]"""
        out = (ctypes.c_int * 4)()
        self._check(self.lib.clbm_diag_contact_angle_slab(self._h, ctypes.c_double(rho_cut), int(base_y_in), out))
        return list(out)

    def interface_heights(self, phi_mid):
        """device-side scans of findInterfaceHeights (PF/apps/rayleighTaylor2D.h:668-708) -> (y at x = 0, y at x = nx/2)"""
        a, b = ctypes.c_int(0), ctypes.c_int(0)
        self._check(self.lib.clbm_diag_interface_heights(self._h, ctypes.c_double(phi_mid), ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    # -- slab exchange
    def halo_buffer(self, phase, side, recv):
        ptr = ctypes.c_void_p()
        nbytes = ctypes.c_size_t(0)
        self._check(self.lib.clbm_halo_buffer(self._h, phase, side, int(recv), ctypes.byref(ptr), ctypes.byref(nbytes)))
        return ptr.value, nbytes.value

    def halo_pack(self, phase):
        self._check(self.lib.clbm_halo_pack(self._h, phase))

    def halo_unpack(self, phase):
        self._check(self.lib.clbm_halo_unpack(self._h, phase))

    def step_stage(self, stage):
        self._check(self.lib.clbm_step_stage(self._h, stage))

    def stream(self):
        return self.lib.clbm_stream(self._h)

    # -- the ring driven from the library (slab_comm.cu)
    def comm_unique_id(self):
        """128 bytes of a fresh ncclUniqueId (rank 0 calls this and broadcasts the bytes)"""
        buf = ctypes.create_string_buffer(128)
        self._check(self.lib.clbm_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, id128, rank, nranks):
        assert len(id128) == 128
        self._check(self.lib.clbm_comm_init(self._h, ctypes.c_char_p(id128), int(rank), int(nranks)))

    def slab_step(self, n=1):
        """n slab steps (stages + both ghost exchanges) inside the library; every rank calls it with the same n"""
        self._check(self.lib.clbm_slab_step(self._h, int(n)))

    PEER_HANDLE_BYTES = 128

    def peer_export(self):
        """opaque handle of this slab's halo mailbox (clbm_peer_export) -> bytes"""
        buf = ctypes.create_string_buffer(self.PEER_HANDLE_BYTES)
        self._check(self.lib.clbm_peer_export(self._h, buf))
        return buf.raw

    def peer_connect(self, left_handle, right_handle):
        """map the ring neighbours' mailboxes (handles from their peer_export, other processes of this node)"""
        self._check(self.lib.clbm_peer_connect(self._h, ctypes.c_char_p(left_handle), ctypes.c_char_p(right_handle)))

    def peer_connect_local(self, left, right):
        """the same for neighbours that are Lattice objects of this process"""
        self._check(self.lib.clbm_peer_connect_local(self._h, left._h, right._h))

    def peer_disconnect(self):
        self._check(self.lib.clbm_peer_disconnect(self._h))

    def ring_kind(self):
        return int(self.lib.clbm_ring_kind(self._h))

    def slab_exchange(self, phase):
        self._check(self.lib.clbm_slab_exchange(self._h, int(phase)))

    def slab_signal(self, phase, boundary=False):
        self._check(self.lib.clbm_slab_signal(self._h, int(phase), 1 if boundary else 0))

    def slab_wait(self, phase, boundary=False):
        self._check(self.lib.clbm_slab_wait(self._h, int(phase), 1 if boundary else 0))

    def overlap_variant(self):
        """0: sequential protocol, 1: interior-first overlap, 2: halo-first overlap (clbm_overlap_variant)"""
        return int(self.lib.clbm_overlap_variant(self._h))

    def overlap_width(self):
        return int(self.lib.clbm_overlap_width(self._h))

    def overlap_supported(self):
        return bool(self.lib.clbm_overlap_supported(self._h))

    def boundary_stream(self):
        return self.lib.clbm_boundary_stream(self._h)


class PinnedArray:
    """numpy view of pinned host memory allocated through the C ABI (clbm_alloc_host)"""

    def __init__(self, n, dtype=np.float64):
        lib = load_library()
        self._lib = lib
        self.nbytes = int(n) * np.dtype(dtype).itemsize
        p = ctypes.c_void_p()
        rc = lib.clbm_alloc_host(self.nbytes, ctypes.byref(p))
        if rc != 0:
            raise ClbmError("clbm_alloc_host(%d bytes) failed: %s" % (self.nbytes, lib.clbm_last_error().decode()))
        self.ptr = p
        buf = (ctypes.c_uint8 * self.nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(n))

    def free(self):
        if self.ptr is not None:
            self.array = None
            self._lib.clbm_free_host(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Pulsatile:
    """Device-resident compliant-vessel case: the counterpart of `LBM_PulsatileBloodFlow2D` and its arrays
    (AB/apps/PulsatileBloodFlow2D.h); `step(n)` is n iterations of the reference loop body :764-790."""

    def __init__(self, N=64, tau=0.75, alpha=0.01, p0_in=0.20, p0_out=0.19, is_severed=1, deformable=1, device=-1):
        self.lib = load_library()
        self.p = P.pulsatile_params(N, tau, alpha, p0_in, p0_out, is_severed, deformable, device)
        h = ctypes.c_void_p()
        self._h = None
        self._check(self.lib.clbm_pulsatile_create(ctypes.byref(self.p), ctypes.byref(h)))
        self._h = h
        v = [ctypes.c_int() for _ in range(3)]
        self._check(self.lib.clbm_pulsatile_info(self._h, ctypes.byref(v[0]), ctypes.byref(v[1]), ctypes.byref(v[2]), None, None))
        self.nx, self.ny, self.tf = (x.value for x in v)
        self.nelem = self.nx * self.ny

    def _check(self, rc):
        if rc != 0:
            raise ClbmError("clbm error %d: %s" % (rc, self.lib.clbm_last_error().decode()))

    def close(self):
        if self._h is not None:
            self.lib.clbm_pulsatile_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, n=1):
        self._check(self.lib.clbm_pulsatile_step(self._h, int(n)))
        return self

    def step_timed(self, n):
        ms = ctypes.c_float()
        self._check(self.lib.clbm_pulsatile_step_timed(self._h, int(n), ctypes.byref(ms)))
        return ms.value

    def sync(self):
        self._check(self.lib.clbm_pulsatile_sync(self._h))

    def launch_count(self):
        return int(self.lib.clbm_pulsatile_launch_count(self._h))

    def kernel_timing_begin(self, cap):
        self._check(self.lib.clbm_pulsatile_kernel_timing_begin(self._h, int(cap)))

    def kernel_timing_end(self):
        ms, n = ctypes.c_float(), ctypes.c_int()
        self._check(self.lib.clbm_pulsatile_kernel_timing_end(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    @property
    def parity(self):
        v = ctypes.c_int()
        self._check(self.lib.clbm_pulsatile_info(self._h, None, None, None, None, ctypes.byref(v)))
        return v.value

    @property
    def t_iter(self):
        v = ctypes.c_int()
        self._check(self.lib.clbm_pulsatile_info(self._h, None, None, None, ctypes.byref(v), None))
        return v.value

    def fields(self, out=None):
        ne = self.nelem
        o = out or {"P": np.empty(ne), "Ux": np.empty(ne), "Uy": np.empty(ne), "flag": np.empty(ne, dtype=np.uint8),
                    "yr1": np.empty(self.nx), "yr2": np.empty(self.nx)}
        self._check(self.lib.clbm_pulsatile_download_fields(self._h, *[_ptr(o.get(k)) for k in ("P", "Ux", "Uy", "flag", "yr1", "yr2")]))
        return o

    def lattice(self):
        a = np.empty(2 * 9 * self.nelem)
        par = ctypes.c_int()
        self._check(self.lib.clbm_pulsatile_download_lattice(self._h, _ptr(a), ctypes.byref(par)))
        return a, par.value

    def upload(self, lattice, flag, Pf, Ux, Uy, yr1, yr2, parity, t_iter):
        self._check(self.lib.clbm_pulsatile_upload(self._h, _ptr(lattice), _ptr(flag), _ptr(Pf), _ptr(Ux), _ptr(Uy), _ptr(yr1),
                                                   _ptr(yr2), int(parity), int(t_iter)))


class YoungLaplace:
    """Device-resident Young-Laplace case: the counterpart of `LBM_Young_Laplace2D` (AB/apps/Young_Laplace2D.h);
    `step(n)` is n iterations of the reference loop body :555-565 (collide_stream_at, parity flip, update_fields)."""

    def __init__(self, nx=128, ny=None, Sigma=0.01, W=4.0, M=0.02, RhoL=0.001, RhoH=1.0, tau=0.8, device=-1):
        self.lib = load_library()
        self.p = P.yl2d_params(nx, ny if ny else nx, Sigma, W, M, RhoL, RhoH, tau, device)
        self.nx, self.ny = self.p.nx, self.p.ny
        self.nelem = self.nx * self.ny
        h = ctypes.c_void_p()
        self._h = None
        self._check(self.lib.clbm_yl2d_create(ctypes.byref(self.p), ctypes.byref(h)))
        self._h = h

    def _check(self, rc):
        if rc != 0:
            raise ClbmError("clbm error %d: %s" % (rc, self.lib.clbm_last_error().decode()))

    def close(self):
        if self._h is not None:
            self.lib.clbm_yl2d_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, n=1):
        self._check(self.lib.clbm_yl2d_step(self._h, int(n)))
        return self

    def step_timed(self, n):
        ms = ctypes.c_float()
        self._check(self.lib.clbm_yl2d_step_timed(self._h, int(n), ctypes.byref(ms)))
        return ms.value

    def sync(self):
        self._check(self.lib.clbm_yl2d_sync(self._h))

    def launch_count(self):
        return int(self.lib.clbm_yl2d_launch_count(self._h))

    def fields(self, out=None):
        o = out or {k: np.empty(self.nelem) for k in ("C", "P", "Rho", "Ux", "Uy")}
        self._check(self.lib.clbm_yl2d_download_fields(self._h, *[_ptr(o.get(k)) for k in ("C", "P", "Rho", "Ux", "Uy")]))
        return o

    def lattice(self):
        a = np.empty(36 * self.nelem)
        par = ctypes.c_int()
        self._check(self.lib.clbm_yl2d_download_lattice(self._h, _ptr(a), ctypes.byref(par)))
        return a, par.value

    def upload(self, lattice, Ux, Uy, parity):
        self._check(self.lib.clbm_yl2d_upload(self._h, _ptr(lattice), _ptr(Ux), _ptr(Uy), int(parity)))

    def reduce(self, kind):
        v = ctypes.c_double()
        self._check(self.lib.clbm_yl2d_reduce(self._h, int(kind), ctypes.byref(v)))
        return v.value
