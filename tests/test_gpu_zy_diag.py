"""Device-side line scans (clbm_diag_contact_angle, clbm_diag_interface_heights) against the reference's serial scans
restated on the downloaded fields -- integers, so the bar is exact equality."""
import numpy as np
import pytest

import _cases

pytestmark = pytest.mark.gpu

pkg = _cases.pkg
P = pkg.params


def contact_angle_scan_ref(rho, flag, nx, ny, rho_cut):
    """SC/apps/contactAngle2D.h:465-505 (calculateContactAngle up to Base / Height), i = y + ny*x"""
    at = lambda x, y: y + ny * x
    base_y = 2
    while base_y < ny and flag[at(0, base_y)] == 0:
        base_y += 1
    if base_y >= ny - 1:
        return base_y, 0, 0
    xmid = nx // 2
    left = right = xmid
    while left > 0 and rho[at(left - 1, base_y)] > rho_cut:
        left -= 1
    while right < nx - 1 and rho[at(right + 1, base_y)] > rho_cut:
        right += 1
    height = 0
    for y in range(base_y, ny):
        if flag[at(xmid, y)] == 0 or not rho[at(xmid, y)] > rho_cut:
            break
        height += 1
    return base_y, max(0, right - left + 1), height


def interface_heights_ref(phi, nx, ny, phi_mid):
    """PF/apps/rayleighTaylor2D.h:668-708: (scan at x = 0 -> `bubble_y`, scan at x = nx/2 -> `spike_y`)"""
    out = []
    for x in (0, nx // 2):
        v = 0
        for y in range(ny - 2, 0, -1):
            if phi[y + ny * x] <= phi_mid:
                v = y
                break
        out.append(v)
    return tuple(out)


@pytest.mark.parametrize("nx,ny,RR,steps", [(96, 48, 14.0, 0), (96, 48, 14.0, 300), (300, 70, 20.0, 150), (64, 300, 9.0, 50)])
def test_contact_angle_scan_matches_serial_scan(nx, ny, RR, steps):
    prm = P.sc_params(P.MODEL_SC_D2Q9, nx, ny, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT)
    with pkg.clbm.Lattice(prm) as lat:
        lat.init_case(P.CASE_SC_CONTACT2D, (0.265, 0.038, RR))
        lat.step(steps)
        got = lat.contact_angle_scan(0.5 * (0.265 + 0.038))
        f, flag = lat.fields(), lat.flags()
    ref = contact_angle_scan_ref(f["s0"], flag, nx, ny, 0.5 * (0.265 + 0.038))
    assert got == ref
    assert ref[1] > 1 and ref[2] > 0       # a droplet is detected


def test_contact_angle_scan_edge_cases():
    """no droplet (uniform gas): base = 1 node run is not formed (rho(xmid +- 1) <= cut), height 0; everything above the cut:
    the run covers the whole row and the column up to the top wall"""
    nx, ny = 40, 24
    prm = P.sc_params(P.MODEL_SC_D2Q9, nx, ny, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT)
    with pkg.clbm.Lattice(prm) as lat:
        lat.init_case(P.CASE_SC_CONTACT2D, (0.265, 0.038, 0.0))     # radius 0: only the centre node (nx/2, 5) is liquid
        f, flag = lat.fields(), lat.flags()
        for cut in (0.15, 0.01, 1.0):
            assert lat.contact_angle_scan(cut) == contact_angle_scan_ref(f["s0"], flag, nx, ny, cut)
        assert lat.contact_angle_scan(0.01) == (2, nx, ny - 3)
    with pytest.raises(pkg.clbm.ClbmError):
        with pkg.clbm.Lattice(P.hcz_params(P.MODEL_HCZ_D2Q9, 16, 66, N=16)) as lat:
            lat.contact_angle_scan(0.1)


@pytest.mark.parametrize("nx,steps", [(32, 0), (32, 400), (200, 100)])
def test_interface_heights_match_serial_scan(nx, steps):
    prm = P.hcz_params(P.MODEL_HCZ_D2Q9, nx, 4 * nx + 2, N=nx)
    mid = 0.5 * (prm.phi_l + prm.phi_g)
    with pkg.clbm.Lattice(prm) as lat:
        lat.init_case(P.CASE_HCZ_RT2D, ())
        lat.step(steps)
        got = lat.interface_heights(mid)
        phi = lat.fields()["s0"]
        none = lat.interface_heights(-1.0)          # nothing is <= -1: both columns report 0
        allhit = lat.interface_heights(10.0)        # everything is: the topmost bulk row
    assert got == interface_heights_ref(phi, prm.nx, prm.ny, mid)
    assert got[0] != got[1] and min(got) > 0        # the cosine perturbation separates spike and bubble
    assert none == (0, 0) and allhit == (prm.ny - 2, prm.ny - 2)


# ---- the drivers that print these diagnostics (C++17 COOLBM binary above the C ABI) ------------------------------
def _coolbm(tmp_path, problem, cfg_name, cfg_text):
    import os
    import subprocess
    apps = os.path.join(_cases.ROOT, "multiphase-lbm_b200", "apps")
    exe = os.path.join(apps, "build", "COOLBM")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", apps])
    cfg = tmp_path / "cfg"
    cfg.mkdir()
    (cfg / cfg_name).write_text(cfg_text)
    r = subprocess.run([exe, problem, str(cfg)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return r.stdout


def test_driver_contact_angle_prints_device_scan(tmp_path):
    out = _coolbm(tmp_path, "contactAngle2D", "config_contactAngle2D.txt",
                  "# skipped\nTT0 0.875\na 1.0\nb 4.0\nR 1.0\nrhow 0.2\nrhol 0.265\nrhog 0.038\nN 48\nulb 0.1\nRe 60\ntau 1.0\n"
                  "max_t 0.6251\nout_freq 100\nvtk_freq 0\nRR 14\n")          # dt = 0.1/48: 300 iterations
    prm = P.sc_params(P.MODEL_SC_D2Q9, 96, 48, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT)
    lines = [l for l in out.split("\n") if l.startswith("Base=")]
    assert len(lines) == 3
    with pkg.clbm.Lattice(prm) as lat:
        lat.init_case(P.CASE_SC_CONTACT2D, (0.265, 0.038, 14.0))
        lat.step(200)
        f, flag = lat.fields(), lat.flags()
    _, base, height = contact_angle_scan_ref(f["s0"], flag, 96, 48, 0.5 * (0.265 + 0.038))
    h, b = float(height), float(base)
    R = (4 * h * h + b * b) / (8 * h)
    theta = np.degrees(np.arctan(0.5 * b / (R - h)))
    theta = theta + 180.0 if theta < 0 else theta
    assert lines[2].startswith("Base=%d Height=%d ContactAngle=" % (base, height))
    assert abs(float(lines[2].split("ContactAngle=")[1].split()[0]) - theta) < 1e-5
    ca = np.loadtxt(tmp_path / "contact_angle.dat")
    assert ca.shape == (3, 3) and ca[2, 0] == base and ca[2, 1] == height


def test_driver_hcz_rayleigh_taylor_writes_device_scan(tmp_path):
    out = _coolbm(tmp_path, "rayleighTaylor2D", "config_rayleighTaylor2D.txt",
                  "# skipped\nRe 3000\nulb 0.04\nN 32\nmax_t 0.2501\nout_freq 100\nvtk_freq 0\nphi_l 0.251\nphi_g 0.024\nrho_l 0.12\n"
                  "rho_g 0.04\na 4\nb 4\nkappa 0.01\ngravity -6.25e-6\n")            # dt = 0.04/32: 200 iterations
    assert "MLUPS" in out or "Throughput" in out
    pos = np.loadtxt(tmp_path / "spike_bubble_position.dat")
    assert pos.shape == (2, 3)
    prm = P.hcz_params(P.MODEL_HCZ_D2Q9, 32, 130, N=32)
    with pkg.clbm.Lattice(prm) as lat:
        lat.init_case(P.CASE_HCZ_RT2D, ())
        lat.step(100)
        phi = lat.fields()["s0"]
    y0, ymid = interface_heights_ref(phi, 32, 130, 0.5 * (prm.phi_l + prm.phi_g))
    dx = 1.0 / 32
    # columns of the file: t, spike (x = nx/2 scan), bubble (x = 0 scan) -- the reference's swapped naming
    assert abs(pos[1, 1] - ymid * dx) < 1e-5 and abs(pos[1, 2] - y0 * dx) < 1e-5


def test_driver_sc_rayleigh_taylor_energy_matches_oracle(tmp_path):
    """COOLBM RayleighTaylor2D (SC/apps/RayleighTaylor2D.h driver surface) against the oracle: energy log, VTK blocks"""
    from _oracle import OracleSim
    out = _coolbm(tmp_path, "RayleighTaylor2D", "config_RayleighTaylor2D.txt",
                  "# skipped\nRe 5.76\nulb 0.04\nN 24\nmax_t 0.3335\nout_freq 100\nvtk_freq 100\nrhol 1.2\nrhog 0.4\ng -5\nrhow 0.2\n"
                  "a 1\nb 4\ngravity -1.25e-5\n")                                       # dt = 0.04/24: 200 iterations
    assert "Rayleigh Taylor 2D problem" in out and "MLUPS" in out
    import os
    assert sorted(f for f in os.listdir(tmp_path) if f.endswith(".vtk")) == ["sol_0000000.vtk", "sol_0000100.vtk"]
    vtk = open(tmp_path / "sol_0000100.vtk").read()
    assert all(t in vtk for t in ("SCALARS Density", "VECTORS Force_ff", "VECTORS Force_fw", "DIMENSIONS 24 98 1"))
    prm = P.sc_rt_params(24, ulb=0.04, N=24, Re=5.76)     # omega = 1 as at the shipped N = 128, Re = 30.72 (1.68 blows up)
    ora = OracleSim(prm).init_case(P.CASE_SC_RT2D, (1.2, 0.4)).step(100)
    u = ora.fields()
    bulk = ora.flag == 1
    dxs, dts = 1.0 / 24, 0.04 / 24
    e_ref = 0.5 * np.sum((u["ux"] ** 2 + u["uy"] ** 2)[bulk]) / (24 * 98) * dxs * dxs / (dts * dts)
    e_drv = np.loadtxt(tmp_path / "energy.dat")[1, 1]
    assert np.isfinite(e_ref) and e_ref > 0
    assert abs(e_drv - e_ref) <= 2e-7 * abs(e_ref)       # energy.dat holds 8 significant digits


@pytest.mark.parametrize("nranks", [2, 3, 5])
def test_scans_on_x_slabs_equal_the_single_slab_scans(nranks):
    """the slab forms of both scans (partial integers per slab in global x, MIN / MAX combined by the caller) against the
    single-slab scans of the same state: the droplet base straddles slab borders, the centre column lies in one slab"""
    slab = pkg.slab
    nx, ny = 120, 48
    prm = P.sc_params(P.MODEL_SC_D2Q9, nx, ny, tau=1.0, rho_w=0.2, sc_force=P.SC_FORCE_CONTACT)
    args = (0.265, 0.038, 22.0)
    cut = 0.5 * (0.265 + 0.038)
    with pkg.clbm.Lattice(prm) as single:
        single.init_case(P.CASE_SC_CONTACT2D, args)
        single.step(60)
        ref = [single.contact_angle_scan(c) for c in (cut, 0.01, 1.0)]
    lats = [pkg.clbm.Lattice(slab.slab_params(prm, r, nranks)) for r in range(nranks)]
    for lat in lats:
        lat.init_case(P.CASE_SC_CONTACT2D, args)
    ring = slab.LocalRing(lats)
    ring.step(60)
    got = [ring.contact_angle_scan(c) for c in (cut, 0.01, 1.0)]
    with pytest.raises(pkg.clbm.ClbmError):
        lats[0].contact_angle_scan(cut)            # the one-call form is for a single slab
    for lat in lats:
        lat.close()
    assert got == ref and ref[0][1] > 30

    prm = P.hcz_params(P.MODEL_HCZ_D2Q9, 60, 242, N=60)
    mid = 0.5 * (prm.phi_l + prm.phi_g)
    with pkg.clbm.Lattice(prm) as single:
        single.init_case(P.CASE_HCZ_RT2D, ())
        single.step(100)
        ref = single.interface_heights(mid)
    lats = [pkg.clbm.Lattice(slab.slab_params(prm, r, nranks)) for r in range(nranks)]
    for lat in lats:
        lat.init_case(P.CASE_HCZ_RT2D, ())
    ring = slab.LocalRing(lats)
    ring.step(100)
    got = ring.interface_heights(mid)
    for lat in lats:
        lat.close()
    assert got == ref and ref[0] > 0 and ref[1] > 0
