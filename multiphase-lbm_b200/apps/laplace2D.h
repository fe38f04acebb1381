// laplace2D.h -- Shan-Chen (Yuan-CS) static droplet, Laplace-law case, on the B200 library.
// Driver surface of SC/apps/laplace2D.h:404-512 (Laplace2D()): same config keys, same derived parameters, same
// terminal / energy.dat / mass.dat / sol_*.vtk output; the lattice lives on the device and the hot line
//     for_each(par_unseq, lattice, lattice + nelem, lbm); *parity = 1 - *parity;
// is clbm_step().
#pragma once
#include <array>

#include "case_common.h"

namespace coolbm {

inline void print_sc_parameters(const char *title, int N, int nx, int ny, double Re, double omega, double ulb, double max_t, double nu)
{
    std::cout << title << "\n"
              << "N      = " << N << '\n' << "nx     = " << nx << '\n' << "ny     = " << ny << '\n'
              << "Re     = " << Re << '\n' << "omega  = " << omega << '\n' << "tau    = " << 1. / omega << '\n'
              << "nu     = " << nu << '\n' << "ulb    = " << ulb << '\n' << "max_t  = " << max_t << '\n';
}

// sol_%07d.vtk with Density, Pressure and the interaction force (SC/apps/laplace2D.h:319-365)
inline void save_vtk_sc(DeviceLattice &lat, int time_iter, double dx)
{
    auto f = lat.fields(true, false);
    const size_t n = lat.nelem();
    std::vector<double> fx(n), fy(n), fz(n);
    check(clbm_download_force(lat.ctx, fx.data(), fy.data(), fz.data()));
    VtkWriter vtk(time_iter, lat.prm.nx, lat.prm.ny, lat.prm.nz, dx);
    vtk.scalars("Density", "float", [&](size_t i) { return f.flag[i] == 0 ? 0.0 : f.s0[i]; });
    vtk.scalars("Pressure", "float", [&](size_t i) { return f.flag[i] == 0 ? 0.0 : f.s1[i]; });
    vtk.vectors("Force", [&](size_t i) { return std::array<double, 3>{fx[i], fy[i], lat.prm.nz > 1 ? fz[i] : 0.0}; });
}

inline void Laplace2D(const std::string &config_dir)
{
    Config cfg{read_config(config_dir + "/config_Laplace2D.txt", "config_Laplace2D.txt")};
    const double Re = cfg.d("Re", 60), ulb = cfg.d("ulb", 0.1), max_t = cfg.d("max_t", 10.0), rhol = cfg.d("rhol", 1.0),
                 rhog = cfg.d("rhog", 0.1), rho_w = cfg.d("rho_w", 0.12), a = cfg.d("a", 1.0), b = cfg.d("b", 4.0), R = cfg.d("R", 1.0),
                 TT0 = cfg.d("TT0", 0.875), gravity = cfg.d("gravity", 0.0), tau_in = cfg.d("tau", -1.0);
    cfg.d("g", 0.0);   // accepted and ignored, as in the reference
    const int N = cfg.i("N", 100), out_freq = cfg.i("out_freq", 400), vtk_freq = cfg.i("vtk_freq", 400);

    double nu, omega, dx = 1.0 / N, dt = dx * ulb;
    if (tau_in > 0.0) { omega = 1.0 / tau_in; nu = (tau_in - 0.5) / 3.0; }
    else { auto p = lb_parameters(ulb, N, Re); nu = p.nu; omega = p.omega; dx = p.dx; dt = p.dt; }
    print_sc_parameters("Laplace 2D problem", N, N, N, Re, omega, ulb, max_t, nu);

    clbm_params prm = default_params(CLBM_MODEL_SC_D2Q9, N, N, 1);
    prm.omega = omega; prm.gravity = gravity; prm.rho_w = rho_w; prm.a = a; prm.b = b; prm.R = R;
    prm.TT = TT0 * (0.3773 * a / (b * R));          // Yuan: TT = TT0 * Tc
    prm.sc_force = CLBM_SC_FORCE_LAPLACE;
    apply_collision_keys(cfg, prm, omega);
    DeviceLattice lat(prm);
    lat.init_case(CLBM_CASE_SC_LAPLACE2D, {rhol, rhog, 10.0});   // iniLattice + inigeom (periodic everywhere)

    Stopwatch sw;
    std::ofstream energyfile("energy.dat"), mass_log("mass.dat");
    double M0 = -1.0;
    const int max_time_iter = static_cast<int>(max_t / dt);
    run_loop(lat, max_time_iter, out_freq, vtk_freq, sw, [&](int time_iter, bool vtk, bool out) {
        if (vtk) save_vtk_sc(lat, time_iter, dx);
        if (!out) return;
        progress_line(time_iter, dt, max_t);
        const double energy = lat.reduce(CLBM_REDUCE_ENERGY) * dx * dx / (dt * dt);
        std::cout << "Average energy: " << std::setprecision(8) << energy << "\n";
        energyfile << std::setw(10) << time_iter * dt << std::setw(16) << std::setprecision(8) << energy << "\n";
        const double M = lat.reduce(CLBM_REDUCE_MASS);
        if (M0 < 0.0) M0 = M;
        std::cout << std::setprecision(12) << "[Mass] M=" << M << "   \xCE\x94M/M0=" << std::setprecision(6) << (M - M0) / M0 * 100.0 << "%\n";
        if (mass_log) mass_log << std::setprecision(16) << time_iter * dt << " " << M << "\n";
    });
    sw.report(lat.nelem());
}

}  // namespace coolbm
