"""The C++17 case drivers (multiphase-lbm_b200/apps, built by __graft_entry__.build()): CPU-side checks of the reference's
error behaviour, and on the GPU the full PulsatileBloodFlow2D run reproducing the reference's shipped VTK files byte
for byte plus a short Laplace2D run checked against the oracle."""
import hashlib
import json
import os
import re
import subprocess

import numpy as np
import pytest

import _cases
from _oracle import OracleSim

APPS = os.path.join(_cases.ROOT, "multiphase-lbm_b200", "apps")
EXE = os.path.join(APPS, "build", "COOLBM")
P = _cases.P


def _exe():
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-s", "-C", APPS])
    return EXE


def test_driver_missing_config_throws_like_reference(tmp_path):
    r = subprocess.run([_exe(), "laplace2D", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1
    assert 'Config file not found. It should be named "config_Laplace2D.txt" in Files_Config.' in r.stderr


def test_driver_unknown_problem():
    r = subprocess.run([_exe(), "noSuchCase"], capture_output=True, text=True)
    assert r.returncode == 2


def test_driver_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([_exe(), "laplace2D", os.path.join(APPS, "Config_Files")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_driver_pulsatile_reproduces_shipped_vtk(tmp_path):
    hashes = json.load(open(os.path.join(_cases.GOLDEN, "pulsatile_vtk_sha256.json")))
    r = subprocess.run([_exe(), "PulsatileBloodFlow2D"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "t=2754 / 2765" in r.stdout and "MLUPS" in r.stdout
    files = sorted(f for f in os.listdir(tmp_path) if f.endswith(".vtk"))
    assert files == sorted(hashes)
    for f in files:
        assert hashlib.sha256(open(tmp_path / f, "rb").read()).hexdigest() == hashes[f], f


@pytest.mark.gpu
def test_driver_laplace2d_matches_oracle(tmp_path):
    cfg = tmp_path / "cfg"
    cfg.mkdir()
    (cfg / "config_Laplace2D.txt").write_text(
        "# first line is skipped\nTT0 0.875\na 1.0\nb 4.0\nR 1.0\nrho_w 0.12\nrhol 0.265\nrhog 0.038\nN 48\nulb 0.01\nRe 6\n"
        "max_t 0.0626\nout_freq 100\nvtk_freq 300\ngravity 0.0\n")
    r = subprocess.run([_exe(), "laplace2D", str(cfg)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "Laplace 2D problem" in r.stdout and re.search(r"result: [0-9.e+-]+ MLUPS", r.stdout)
    # dt = ulb / N -> max_time_iter = int(0.0626 * 4800) = 300 steps; the VTK of iteration 300 is not written (loop ends at 299)
    assert sorted(f for f in os.listdir(tmp_path) if f.endswith(".vtk")) == ["sol_0000000.vtk"]
    mass = np.loadtxt(tmp_path / "mass.dat")
    assert mass.shape == (3, 2) and abs(mass[-1, 1] - mass[0, 1]) / mass[0, 1] < 1e-12
    prm = P.sc_params(P.MODEL_SC_D2Q9, 48, 48, ulb=0.01, N=48, Re=6.0)
    ora = OracleSim(prm).init_case(P.CASE_SC_LAPLACE2D, (0.265, 0.038, 10.0))
    ora.step(200)
    u = ora.fields()
    e_ref = 0.5 * np.sum(u["ux"] ** 2 + u["uy"] ** 2) / (48 * 48) * (1 / 48.) ** 2 / (0.01 / 48.) ** 2
    e_drv = np.loadtxt(tmp_path / "energy.dat")[2, 1]
    assert abs(e_drv - e_ref) <= 2e-8 * abs(e_ref) + 1e-30     # energy.dat holds 8 significant digits


@pytest.mark.gpu
def test_driver_two_layered_flow_matches_oracle(tmp_path):
    """COOLBM twoLayeredFlow2D (the SC reference's default problem) against the oracle: mass log and energy log"""
    cfg = tmp_path / "cfg"
    cfg.mkdir()
    (cfg / "config_twoLayeredFlow2D.txt").write_text(
        "N 40\nulb 0.1\nRe 60\nmax_t 1.0001   # 400 steps\nout_freq 200\nvtk_freq 200\na 1.0\nb 4.0\nR 1.0\nTT0 0.95\nrhol 0.21\n"
        "rhog 0.067\nrho_w 0.067\nh_lower 0.30\nw_int 4\ngx 1e-6\ngy 0.0\n")
    r = subprocess.run([_exe(), "twoLayeredFlow2D", str(cfg)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "Two-Layered Flow 2-D" in r.stdout and "p_shift = " in r.stdout and "MLUPS" in r.stdout
    assert sorted(f for f in os.listdir(tmp_path) if f.endswith(".vtk")) == ["sol_0000000.vtk", "sol_0000200.vtk"]
    vtk = open(tmp_path / "sol_0000200.vtk").read()
    assert all(tag in vtk for tag in ("SCALARS Density", "SCALARS Pressure", "VECTORS Force", "VECTORS Velocity"))
    prm = P.sc_layered_params(10, 41, ulb=0.1, N=40, Re=60.0, gx=1e-6)
    ora = OracleSim(prm).init_case(P.CASE_SC_LAYERED2D, (0.21, 0.067, 0.30, 4.0))
    ora.step(200)
    u = ora.fields()
    bulk = ora.flag == 1
    e_ref = 0.5 * np.sum((u["ux"] ** 2 + u["uy"] ** 2)[bulk]) / (10 * 41) * (1 / 40.) ** 2 / (0.1 / 40.) ** 2
    e_drv = np.loadtxt(tmp_path / "energy.dat")[1, 1]
    assert abs(e_drv - e_ref) <= 2e-9 * abs(e_ref) + 1e-30
    m_drv = np.loadtxt(tmp_path / "mass.dat")[1, 1]
    assert abs(m_drv - np.sum(u["s0"][bulk])) <= 1e-12 * m_drv


@pytest.mark.gpu
def test_driver_young_laplace_matches_reference_writer(tmp_path):
    """COOLBM Young_Laplace2D (the AB reference's default problem): logs against the oracle, VTK layout of AB:374-421"""
    from _oracle import YL2DOracle
    cfg = tmp_path / "cfg"
    cfg.mkdir()
    (cfg / "config_laplace2D.txt").write_text("# test\nN 48\ntf 200\nout_freq 100\nvtk_freq 200\nSigma 0.01\nW 4.0\nM 0.02\nRhoL 0.001\nRhoH 1.0\ntau 0.8\n")
    r = subprocess.run([_exe(), "Young_Laplace2D", str(cfg)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "it =     200   [100.0%]" in r.stdout and "Throughput:" in r.stdout
    assert sorted(f for f in os.listdir(tmp_path) if f.endswith(".vtk")) == ["sol_0000000.vtk", "sol_0000200.vtk"]
    o = YL2DOracle(48, 48).step(200)
    f = o.fields()
    mass = np.loadtxt(tmp_path / "mass.dat")
    assert mass.shape == (3, 2) and abs(mass[2, 1] - f["Rho"].sum()) < 1e-11 * f["Rho"].sum()
    e_ref = 0.5 * np.sum(f["Ux"] ** 2 + f["Uy"] ** 2) / (48 * 48)
    assert abs(np.loadtxt(tmp_path / "energy.dat")[2, 1] - e_ref) <= 2e-8 * e_ref
    # the phi block of the VTK file: 48 rows of 48 floats printed like `os << float(x)`
    lines = open(tmp_path / "sol_0000200.vtk").read().split("\n")
    k = lines.index("SCALARS phi float 1")
    row0 = np.array(lines[k + 2].split(), dtype=np.float64)
    ref_row0 = np.array([float("%g" % np.float32(f["C"][0 + 48 * x])) for x in range(48)])
    np.testing.assert_array_equal(row0, ref_row0)
    o.close()


@pytest.mark.parametrize("problem,expect", [
    ("laplace2D", ["N      = 256", "omega  = 0.561798", "max_t  = 20.0001"]),
    ("contactAngle2D", ["nx     = 400", "ny     = 200", "rho_w=0.2"]),
    ("twoLayeredFlow2D", ["ny     = 101", "TT0 (reduced)=0.95", "rho_w=0.067", "p_shift = "]),
    ("droplet3D", ["N      = 256", "tau    = 1"]),
    ("rayleighTaylor2D", ["ny     = 1026", "omega  = 1.95986"]),
    ("RayleighTaylor2D", ["Rayleigh Taylor 2D problem", "ny     = 514", "omega  = 1\n", "nu     = 0.166667"]),
    ("twoLayeredPF2D", ["ny     = 101", "w_int   = 2", "Gx_const= 1e-08"]),
    ("laplace3D", ["nz     = 128", "omega  = 0.877193"]),
])
def test_driver_parses_shipped_config(problem, expect):
    """every shipped Config_Files/*.txt goes through its driver's reader (first-line quirk, trailing comments) -- checked on
    the parameter banner the drivers print before they touch the device"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("would run the full case on a GPU box")
    r = subprocess.run([_exe(), problem, os.path.join(APPS, "Config_Files")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "no CUDA device" in r.stderr
    for e in expect:
        assert e in r.stdout, (problem, e, r.stdout)


def test_driver_accepts_the_mrt_keys(tmp_path):
    """`collision MRT` + rates are optional keys of this library's HCZ D2Q9 drivers (absent from the reference's files); checked
    on the banner printed before the device is touched"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("would run the case on a GPU box (covered by tests/test_gpu_zx_hcz_mrt.py)")
    cfg = tmp_path / "cfg"
    cfg.mkdir()
    (cfg / "config_rayleighTaylor2D.txt").write_text(
        "# skipped\nRe 3000\nulb 0.04\nN 32\nmax_t 0.1\nout_freq 100\nvtk_freq 0\nphi_l 0.251\nphi_g 0.024\nrho_l 0.12\nrho_g 0.04\n"
        "a 4\nb 4\nkappa 0.01\ngravity -6.25e-6\ncollision MRT\ns_e 1.8\n")
    r = subprocess.run([_exe(), "rayleighTaylor2D", str(cfg)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "no CUDA device" in r.stderr
    assert "collision = MRT  s_e = 1.8  s_eps = 1.99489  s_q = 1.99489" in r.stdout
    assert "unknown parameter" not in r.stderr
