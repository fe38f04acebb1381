// twoLayeredPF2D.h -- HCZ phase-field two-layered channel flow (10 x (N+1), walls y = 0, N, x body force) on the B200
// library.  Driver surface of PF/apps/twoLayeredFlow2D.h:759-905 (twoLayered2D()): config_twoLayeredFlow2D.txt keys, energy.dat,
// mass.dat, density_probe.dat header, sol_*.vtk with phi, density, Velocity, Flag (:640-700).
#pragma once
#include <array>

#include "rayleighTaylor2D.h"
#include "twoLayeredFlow2D.h"

namespace coolbm {

inline void twoLayeredPF2D(const std::string &config_dir)
{
    Config cfg{read_config_lines(config_dir + "/config_twoLayeredPF2D.txt",
                                 "Config file not found. It should be named \"config_twoLayeredFlow2D.txt\" in Files_Config.")};
    const double Re = cfg.d("Re", 60), ulb = cfg.d("ulb", 0.1), max_t = cfg.d("max_t", 10.0), phi_l = cfg.d("phi_l", 0.25),
                 phi_g = cfg.d("phi_g", 0.02), rho_l = cfg.d("rho_l", 1.0), rho_g = cfg.d("rho_g", 0.1), a = cfg.d("a", 4.0),
                 b = cfg.d("b", 4.0), kappa = cfg.d("kappa", 0.01), h_lower = cfg.d("h_lower", 0.5), gx = cfg.d("gx", 0.0),
                 Gx_const = cfg.d("Gx_const", 0.0), tau_in = cfg.d("tau", -1.0);
    const int N = cfg.i("N", 100), out_freq = cfg.i("out_freq", 400), vtk_freq = cfg.i("vtk_freq", 400), w_int = cfg.i("w_int", 4);
    cfg.i("data_freq", 0);
    const int nx = 10, ny = N + 1;
    double nu, omega, dx = 1.0 / N, dt = dx * ulb;
    if (tau_in > 0.0) { omega = 1.0 / tau_in; nu = (tau_in - 0.5) / 3.0; }
    else { auto p = lb_parameters(ulb, N, Re); nu = p.nu; omega = p.omega; dx = p.dx; dt = p.dt; }
    print_hcz_parameters("Two-layered flow 2D problem (phase field)", N, nx, ny, 1, Re, omega, ulb, max_t, nu);
    std::cout << "h_lower = " << h_lower << "\nw_int   = " << w_int << "\ngx      = " << gx << "\nGx_const= " << Gx_const << "\n";

    clbm_params prm = default_params(CLBM_MODEL_HCZ_D2Q9, nx, ny, 1);
    prm.omega = omega; prm.phi_l = phi_l; prm.phi_g = phi_g; prm.rho_l = rho_l; prm.rho_g = rho_g; prm.a = a; prm.b = b; prm.kappa = kappa;
    prm.sc_force = CLBM_HCZ_FORCE_LAYERED; prm.gx = gx; prm.gx_const = Gx_const;
    apply_collision_keys(cfg, prm, omega);
    DeviceLattice lat(prm);
    lat.init_case(CLBM_CASE_HCZ_LAYERED2D, {h_lower, (double)w_int});

    Stopwatch sw;
    std::ofstream efile("energy.dat"), mass_log("mass.dat"), dprobe("density_probe.dat");
    dprobe << "# t  rho_center  rho_qbot  rho_qtop\n";
    double M0 = -1.0;
    run_loop(lat, static_cast<int>(max_t / dt), out_freq, vtk_freq, sw, [&](int time_iter, bool vtk, bool out) {
        if (vtk) {
            auto f = lat.fields(false, true);
            VtkWriter w(time_iter, nx, ny, 1, dx);
            w.scalars("phi", "float", [&](size_t i) { return f.s0[i]; });
            w.scalars("density", "float", [&](size_t i) { return f.s2[i]; });
            w.vectors("Velocity", [&](size_t i) { return std::array<double, 3>{f.flag[i] == 0 ? 0.0 : f.ux[i], f.flag[i] == 0 ? 0.0 : f.uy[i], 0.0}; });
            w.scalars("Flag", "int", [&](size_t i) { return f.flag[i] == 0 ? 1 : 0; });
        }
        if (!out) return;
        progress_line(time_iter, dt, max_t);
        const double energy = lat.reduce(CLBM_REDUCE_ENERGY) * dx * dx / (dt * dt);
        std::cout << "Average energy: " << std::setprecision(10) << energy << "\n";
        efile << std::setw(10) << time_iter * dt << std::setw(16) << std::setprecision(10) << energy << "\n";
        const double M = lat.reduce(CLBM_REDUCE_MASS);
        if (M0 < 0.0) M0 = M;
        std::cout << std::setprecision(12) << "[Mass] M=" << M << "   \xCE\x94M/M0=" << std::setprecision(6) << (M - M0) / M0 * 100.0 << "%\n";
        if (mass_log) mass_log << std::setprecision(16) << time_iter * dt << " " << M << "\n";
    });
    sw.report(lat.nelem());
}

}  // namespace coolbm
