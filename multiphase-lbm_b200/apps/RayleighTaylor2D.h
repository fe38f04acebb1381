// RayleighTaylor2D.h -- Shan-Chen Rayleigh-Taylor instability (psi = 1 - exp(-rho), Guo forcing) on the B200 library.
// Driver surface of SC/apps/RayleighTaylor2D.h:577-670 (RayleighTaylor2D(); commented out of the shipped COOLBM.cpp :74,
// served here as problem "RayleighTaylor2D" -- capital R; "rayleighTaylor2D" is the HCZ phase-field case):
// config_RayleighTaylor2D.txt through the older reader (first line skipped), N x (4N + 2) lattice with walls y = 0, ny-1,
// tanh interface of width 2.5 at ny/2 + 0.1 nx cos(2 pi x / (nx - 1)), energy.dat from u_eq = u + F/(2 rho) (:503-516),
// sol_*.vtk with Density, Force_ff and Force_fw (:439-500; force_fw is multiplied by 0 in the reference, so it is all zeros).
// One deviation, stated: the reference's VTK evaluates force_ff on the wall rows too, where it indexes flag[] out of
// bounds (y - 1 = -1, :249-251); this driver writes 0 there.
#pragma once
#include <array>
#include <cmath>

#include "laplace2D.h"

namespace coolbm {

inline void RayleighTaylor2D(const std::string &config_dir)
{
    Config cfg{read_config(config_dir + "/config_RayleighTaylor2D.txt", "config_rayleighTaylor2D.txt")};
    const double Re = cfg.d("Re", 0), ulb = cfg.d("ulb", 0), max_t = cfg.d("max_t", 0), rhol = cfg.d("rhol", 0), rhog = cfg.d("rhog", 0),
                 rhow = cfg.d("rhow", 0), g = cfg.d("g", 0), a = cfg.d("a", 0), b = cfg.d("b", 0), gravity = cfg.d("gravity", 0);
    const int N = cfg.i("N", 0), out_freq = cfg.i("out_freq", 0), vtk_freq = cfg.i("vtk_freq", 0);
    const bool has_collision = cfg.has("collision");
    const int nx = N, ny = 4 * N + 2;

    const auto lb = lb_parameters(ulb, N, Re);
    const double dx = lb.dx, dt = lb.dt;
    std::cout << "Rayleigh Taylor 2D problem\n"
              << "N      = " << N << '\n' << "nx     = " << nx << '\n' << "ny     = " << ny << '\n' << "Re     = " << Re << '\n'
              << "omega  = " << lb.omega << '\n' << "tau    = " << 1. / lb.omega << '\n' << "nu     = " << lb.nu << '\n'
              << "ulb    = " << ulb << '\n' << "max_t  = " << max_t << '\n';

    clbm_params prm = default_params(CLBM_MODEL_SC_D2Q9, nx, ny, 1);
    prm.omega = lb.omega; prm.gravity = gravity; prm.rho_w = rhow; prm.a = a; prm.b = b; prm.G = g;
    prm.sc_force = CLBM_SC_FORCE_EXPGUO;
    if (has_collision) apply_collision_keys(cfg, prm, lb.omega);   // optional keys of this library
    cfg.warn_unknown();
    DeviceLattice lat(prm);
    lat.init_case(CLBM_CASE_SC_RT2D, {rhol, rhog});

    Stopwatch sw;
    std::ofstream energyfile("energy.dat");
    run_loop(lat, static_cast<int>(max_t / dt), out_freq, vtk_freq, sw, [&](int time_iter, bool vtk, bool out) {
        if (vtk) {
            auto f = lat.fields(false, false);
            const size_t n = lat.nelem();
            std::vector<double> fx(n), fy(n);
            check(clbm_download_force(lat.ctx, fx.data(), fy.data(), nullptr));
            VtkWriter w(time_iter, nx, ny, 1, 1. / nx);
            w.scalars("Density", "float", [&](size_t i) { return f.flag[i] == 0 ? 0.0 : f.s0[i]; });
            w.vectors("Force_ff", [&](size_t i) { return std::array<double, 3>{fx[i], fy[i], 0.0}; });
            w.vectors("Force_fw", [&](size_t) { return std::array<double, 3>{0.0, 0.0, 0.0}; });
        }
        if (!out) return;
        std::cout << "Saving profiles at iteration " << time_iter << ", t = " << std::setprecision(4) << time_iter * dt
                  << std::setprecision(3) << " [" << time_iter * dt / max_t * 100. << "%]" << std::endl;
        const double energy = lat.reduce(CLBM_REDUCE_ENERGY) * dx * dx / (dt * dt);
        std::cout << "Average energy: " << std::setprecision(8) << energy << std::endl;
        energyfile << std::setw(10) << time_iter * dt << std::setw(16) << std::setprecision(8) << energy << std::endl;
    });
    sw.report(lat.nelem());
}

}  // namespace coolbm
