"""The partial-sum scheme of the single-sweep HCZ D3Q19 kernel on the CPU (no GPU needed).

csrc/hcz3d_edges.cuh holds the index maps and the arithmetic the kernel uses to build the moments of the NEXT step while it
pushes (per-tile gathers, A/B/C groups along x, edge arrays for the tile perimeter, the periodic wrap of the march).  They are
`__host__ __device__` under CLBM_HOST_CHECK; tests/host_check/hcz3d_edges_host.cu runs them tile by tile in the kernel's order.
Fed with random post-collision populations, the assembled phi, P_term, jx, jy, jz must equal a direct periodic gather
    M(x) = sum_k post_k(x - c_k)
to round-off.  This pins the scheme (who writes which slot, who adds what, the wrap); the kernel's own pipeline is checked
against the oracle on the GPU."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import _cases

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_check", "hcz3d_edges_host.cu")
LIB = os.path.join(HERE, "host_check", "_build", "libhcz3d_edges_host.so")
CSRC = os.path.join(_cases.ROOT, "multiphase-lbm_b200", "csrc")

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists(LIB), reason="nvcc not available")

C19 = [(-1, 0, 0), (0, -1, 0), (0, 0, -1), (-1, -1, 0), (-1, 1, 0), (-1, 0, -1), (-1, 0, 1), (0, -1, -1), (0, -1, 1)]
C19 = C19 + [(0, 0, 0)] + [tuple(-v for v in c) for c in C19]


def _lib():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("hcz3d_edges.cuh", "lattice.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["nvcc", "-std=c++17", "-O2", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared", "-o", LIB, SRC])
    L = ctypes.CDLL(LIB)
    L.host_check_hcz3d_edges.restype = ctypes.c_int
    return L


def direct(post, wrap):
    """[5][nx][ny][nz]: phi, P_term, jx, jy, jz of the populations that arrive (push streaming, periodic y, z; x periodic or open)"""
    f, g = post[:19], post[19:]
    out = np.zeros((5,) + f.shape[1:])
    for k, (cx, cy, cz) in enumerate(C19):
        sf = np.roll(f[k], (cx, cy, cz), axis=(0, 1, 2))
        sg = np.roll(g[k], (cx, cy, cz), axis=(0, 1, 2))
        if not wrap:          # what would have wrapped around in x never arrives
            if cx > 0:
                sf[0] = 0.0; sg[0] = 0.0
            elif cx < 0:
                sf[-1] = 0.0; sg[-1] = 0.0
        out[0] += sf
        out[1] += sg
        out[2] += cx * sg
        out[3] += cy * sg
        out[4] += cz * sg
    return out


@pytest.mark.parametrize("ty,tz,nx,ny,nz", [(8, 32, 5, 16, 64), (8, 32, 4, 8, 32), (4, 8, 7, 12, 24), (4, 8, 3, 4, 8)])
@pytest.mark.parametrize("wrap", [1, 0])
def test_tile_partial_sums_assemble_to_the_direct_moments(ty, tz, nx, ny, nz, wrap):
    rng = np.random.default_rng(ty * 1000 + nx * 10 + wrap)
    post = rng.random((38, nx, ny, nz))
    out = np.zeros((5, nx, ny, nz))
    dp = ctypes.POINTER(ctypes.c_double)
    rc = _lib().host_check_hcz3d_edges(ty, tz, nx, ny, nz, post.ctypes.data_as(dp), out.ctypes.data_as(dp), wrap)
    assert rc == 0
    ref = direct(post, wrap)
    sel = slice(None) if wrap else slice(1, nx - 1)      # open x: the two boundary planes are rebuilt from the populations by the slab code
    err = np.max(np.abs(out[:, sel] - ref[:, sel]))
    assert err < 5e-14, err


def test_every_edge_slot_has_exactly_one_writer_and_one_reader():
    """the index maps alone: over all tiles, the ring slots are distinct and are exactly the slots the border nodes read"""
    lib = _lib()   # noqa: F841  (build check)
    ty, tz, ny, nz = 8, 32, 24, 96
    nTY, nTZ = ny // ty, nz // tz
    off = {}
    off["eyb"] = 0
    off["eyt"] = off["eyb"] + nTY * nz
    off["ezl"] = off["eyt"] + nTY * nz
    off["ezr"] = off["ezl"] + ny * nTZ
    off["ec"] = off["ezr"] + ny * nTZ
    eplane = off["ec"] + 4 * nTY * nTZ

    def edge_offsets(yy, zz):
        ly, lz, R, C = yy % ty, zz % tz, yy // ty, zz // tz
        yb, yt, zl, zr = ly == 0, ly == ty - 1, lz == 0, lz == tz - 1
        e0 = off["eyb"] + R * nz + zz if yb else (off["eyt"] + R * nz + zz if yt else -1)
        e1 = off["ezl"] + yy * nTZ + C if zl else (off["ezr"] + yy * nTZ + C if zr else -1)
        e2 = off["ec"] + (((2 if yt else 0) + (1 if zr else 0)) * nTY + R) * nTZ + C if (yb or yt) and (zl or zr) else -1
        return e0, e1, e2

    readers = set()
    for yy in range(ny):
        for zz in range(nz):
            readers.update(e for e in edge_offsets(yy, zz) if e >= 0)
    writers = []
    for y0 in range(0, ny, ty):
        for z0 in range(0, nz, tz):
            for dy in range(-1, ty + 1):
                for dz in range(-1, tz + 1):
                    ys, zs = dy < 0 or dy >= ty, dz < 0 or dz >= tz
                    if not (ys or zs):
                        continue
                    e = edge_offsets((y0 + dy) % ny, (z0 + dz) % nz)
                    writers.append(e[2] if (ys and zs) else (e[0] if ys else e[1]))
    assert all(w >= 0 for w in writers)
    assert len(set(writers)) == len(writers)          # one writer per slot
    assert set(writers) == readers                    # every slot that is read is written, and nothing else
    assert len(readers) == eplane                     # the arrays have no unused slot


@pytest.mark.parametrize("ty,tz,ny,nz", [(8, 32, 40, 64), (8, 32, 16, 32), (8, 32, 8, 32), (4, 8, 20, 24)])
def test_box_stores_plus_thread_stores_are_the_periodic_push(ty, tz, ny, nz):
    """hcz3d_sweep.cu, VAR bit 2: the c_z = 0 directions of the tiles away from the first / last tile row leave as TMA box stores,
    everything else as thread-level stores with the periodic wrap.  With the rule the kernel uses (push_by_box / push_box_start):
    no box breaks the device's rule for box stores (non-negative, 16-byte aligned start) or is clipped, every (node, direction)
    slot is written exactly once, and the result is the periodic push."""
    rng = np.random.default_rng(ny * 100 + nz)
    post = rng.random((19, ny, nz))
    out = np.full((19, ny, nz), np.nan)
    writes = np.zeros((19, ny, nz), dtype=np.int32)
    dp = ctypes.POINTER(ctypes.c_double)
    L = _lib()
    rc = L.host_check_hcz3d_box_push(ty, tz, ny, nz, post.ctypes.data_as(dp), out.ctypes.data_as(dp), writes.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    assert rc == 0
    assert np.all(writes == 1)
    for k, (cx, cy, cz) in enumerate(C19):
        assert np.array_equal(out[k], np.roll(post[k], (cy, cz), axis=(0, 1))), k
    # the boxes are really used: 9 directions with c_z = 0 in every tile of the interior tile rows
    rows = max(ny // ty - 2, 0)
    assert L.host_check_hcz3d_box_count(ty, tz, ny, nz) == 9 * rows * (nz // tz)
