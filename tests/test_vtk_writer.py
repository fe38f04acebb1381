"""coolbm::VtkWriter (multiphase-lbm_b200/apps/case_common.h): the legacy ASCII layout of the reference writers and the
binary XML ImageData alternative (COOLBM_VTK_FORMAT=vti) hold the same numbers.  No device involved."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import _cases

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_check", "vtk_writer_check.cpp")
EXE = os.path.join(HERE, "host_check", "_build", "vtk_writer_check")
HDR = os.path.join(_cases.ROOT, "multiphase-lbm_b200", "apps", "case_common.h")


def _exe():
    if not os.path.exists(EXE) or max(os.path.getmtime(SRC), os.path.getmtime(HDR)) > os.path.getmtime(EXE):
        os.makedirs(os.path.dirname(EXE), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O1", SRC, "-o", EXE])
    return EXE


def _expected(nx, ny, nz):
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")     # VTK point order: x fastest
    dens = (1.0 + 0.5 * x + 0.25 * y + 0.125 * z).astype(np.float32)
    flag = ((y == 0) | (y == ny - 1)).astype(np.int32)
    force = np.stack([1.0 * x, -1.0 * y, 0.5 * z], axis=-1).astype(np.float32)
    vel = np.stack([0.1 * x, 0.2 * y, 0.0 * z], axis=-1).astype(np.float32)
    return dens, flag, force, vel


def _read_vti(path):
    raw = open(path, "rb").read()
    head, rest = raw.split(b'<AppendedData encoding="raw">', 1)
    data = rest[rest.index(b"_") + 1:]
    arrays = {}
    for m in re.finditer(rb'<DataArray type="(\w+)" Name="(\w+)"(?: NumberOfComponents="(\d+)")? format="appended" offset="(\d+)"/>', head):
        typ, name, ncomp, off = m.group(1).decode(), m.group(2).decode(), int(m.group(3) or 1), int(m.group(4))
        nbytes, = struct.unpack_from("<Q", data, off)
        dt = {"Float32": np.float32, "Int32": np.int32}[typ]
        a = np.frombuffer(data, dtype=dt, count=nbytes // 4, offset=off + 8)
        arrays[name] = a.reshape(-1, ncomp) if ncomp > 1 else a
    return head.decode(), arrays


@pytest.mark.parametrize("dims", [(5, 4, 1), (4, 3, 6)])
def test_vti_blocks_hold_the_fields(tmp_path, dims):
    nx, ny, nz = dims
    subprocess.check_call([_exe()] + [str(d) for d in dims], cwd=tmp_path, env=dict(os.environ, COOLBM_VTK_FORMAT="vti"))
    assert os.listdir(tmp_path) == ["sol_0000042.vti"]
    head, arr = _read_vti(tmp_path / "sol_0000042.vti")
    assert 'WholeExtent="0 %d 0 %d 0 %d"' % (nx - 1, ny - 1, nz - 1) in head and 'Spacing="0.125 0.125 0.125"' in head
    assert 'byte_order="LittleEndian" header_type="UInt64"' in head
    dens, flag, force, vel = _expected(nx, ny, nz)
    np.testing.assert_array_equal(arr["Density"], dens.ravel())
    np.testing.assert_array_equal(arr["Flag"], flag.ravel())
    np.testing.assert_array_equal(arr["Force"], force.reshape(-1, 3))
    np.testing.assert_array_equal(arr["Velocity"], vel.reshape(-1, 3))


def test_legacy_ascii_layout_is_the_reference_writers(tmp_path):
    """default format: SC/apps/laplace2D.h:319-365 (header, y-outer / x-inner rows, trailing blank), PF Flag block with the
    blank line after the plane (PF/apps/rayleighTaylor2D.h:763-780), AB 'u v 0' rows (AB/apps/Young_Laplace2D.h:406-412)"""
    nx, ny = 5, 4
    env = {k: v for k, v in os.environ.items() if k != "COOLBM_VTK_FORMAT"}
    subprocess.check_call([_exe(), str(nx), str(ny), "1"], cwd=tmp_path, env=env)
    assert os.listdir(tmp_path) == ["sol_0000042.vtk"]
    lines = open(tmp_path / "sol_0000042.vtk").read().split("\n")
    assert lines[:10] == ["# vtk DataFile Version 2.0", "iteration 42", "ASCII", "", "DATASET STRUCTURED_POINTS", "DIMENSIONS 5 4 1",
                          "ORIGIN 0 0 0", "SPACING 0.125 0.125 0.125", "", "POINT_DATA 20"]
    assert lines[10:12] == ["SCALARS Density float 1", "LOOKUP_TABLE default"]
    dens, flag, force, vel = _expected(nx, ny, 1)
    rows = np.array([l.split() for l in lines[12:12 + ny]], dtype=np.float64)
    np.testing.assert_array_equal(rows, dens[0].astype(np.float64))
    k = lines.index("SCALARS Flag int 1")
    assert [l.strip() for l in lines[k + 2:k + 2 + ny]] == [" ".join(str(v) for v in r) for r in flag[0]] and lines[k + 2 + ny] == ""
    k = lines.index("VECTORS Force float")
    got = np.array([l.split() for l in lines[k + 1:k + 1 + nx * ny]], dtype=np.float64)
    np.testing.assert_array_equal(got, force[0].reshape(-1, 3).astype(np.float64))
    k = lines.index("VECTORS Velocity float")
    assert lines[k + 1] == "0 0 0" and lines[k + 1 + nx] == ""          # a blank line after every row of nx vectors
