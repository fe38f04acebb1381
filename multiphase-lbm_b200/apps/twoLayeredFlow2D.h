// twoLayeredFlow2D.h -- Shan-Chen (Yuan-CS EOS, constant coupling G) two-layered channel flow on the B200 library: the
// problem the reference's SC build runs by default (SC/apps/COOLBM.cpp:68).  Driver surface of
// SC/apps/twoLayeredFlow2D.h:457-598 (twoLayeredFlow2D()): per-line `key value # comment` config (this driver's reader
// parses every line, including the first), 10 x (N+1) lattice, p_shift from the 601-point scan (:535-546), mass.dat /
// energy.dat / sol_*.vtk with Density, Pressure (EOS), Force and Velocity (:349-410).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>

#include "laplace2D.h"

namespace coolbm {

// unlike the older drivers, this one reads line by line through an istringstream (:478-482): no first-line quirk
inline std::map<std::string, std::string> read_config_lines(const std::string &path, const std::string &msg)
{
    std::ifstream in(path);
    if (!in.is_open()) throw std::invalid_argument(msg);
    std::map<std::string, std::string> kv;
    std::string line, param, value;
    while (std::getline(in, line)) {
        if (auto pos = line.find('#'); pos != std::string::npos) line.erase(pos);
        std::istringstream iss(line);
        if (!(iss >> param >> value)) continue;
        kv[param] = value;
    }
    return kv;
}

inline void twoLayeredFlow2D(const std::string &config_dir)
{
    Config cfg{read_config_lines(config_dir + "/config_twoLayeredFlow2D.txt",
                                 "Config file not found. Expected: config_twoLayeredFlow2D.txt in current directory.")};
    const double Re = cfg.d("Re", 60), ulb = cfg.d("ulb", 0.1), max_t = cfg.d("max_t", 10.0), rhol = cfg.d("rhol", 1.0),
                 rhog = cfg.d("rhog", 0.1), a = cfg.d("a", 1.0), b = cfg.d("b", 4.0), R = cfg.d("R", 1.0), TT0 = cfg.d("TT0", 0.875),
                 tau_in = cfg.d("tau", -1.0), h_lower = cfg.d("h_lower", 0.3), gx = cfg.d("gx", 0.0), gy = cfg.d("gy", 0.0),
                 G = cfg.d("G", -1.0);
    const double rho_w = cfg.has("rhow") ? cfg.d("rhow", 0.12) : cfg.d("rho_w", 0.12);
    cfg.d("Gx_const", 0.0);   // accepted, unused (as in the reference)
    const int N = cfg.i("N", 100), out_freq = cfg.i("out_freq", 400), vtk_freq = cfg.i("vtk_freq", 400), w_int = cfg.i("w_int", 4);
    const int nx = 10, ny = N + 1;

    auto lb = lb_parameters(ulb, N, Re);
    double omega = lb.omega;
    const double nu = lb.nu, dx = lb.dx, dt = lb.dt;
    if (tau_in > 0.0) omega = 1.0 / tau_in;
    std::cout << "Two-Layered Flow 2-D (Yuan\xE2\x80\x93" "CS)\n"
              << "N      = " << N << '\n' << "nx     = " << nx << '\n' << "ny     = " << ny << '\n' << "Re     = " << Re << '\n'
              << "omega  = " << omega << '\n' << "tau    = " << 1. / omega << '\n' << "nu     = " << nu << '\n'
              << "ulb    = " << ulb << '\n' << "max_t  = " << max_t << '\n' << "h_lower (frac of H) = " << h_lower << '\n'
              << "w_int (nodes) = " << w_int << '\n' << "G (coupling) = " << G << '\n';

    clbm_params prm = default_params(CLBM_MODEL_SC_D2Q9, nx, ny, 1);
    const double Tc = 0.3773 * a / (b * R);
    prm.omega = omega; prm.rho_w = rho_w; prm.a = a; prm.b = b; prm.R = R; prm.TT = TT0 * Tc;
    prm.sc_force = CLBM_SC_FORCE_CONSTG; prm.gx = gx; prm.gy = gy; prm.G = G;
    apply_collision_keys(cfg, prm, omega);
    std::cout << std::setprecision(6) << "CS params: a=" << a << " b=" << b << " R=" << R << "\n"
              << "TT0 (reduced)=" << TT0 << "  Tc=" << Tc << "  TT (abs)=" << prm.TT << "\n"
              << "rho_l=" << rhol << "  rho_g=" << rhog << "  rho_w=" << rho_w << "\n";
    // p_shift so that S(rho) = cs2 rho - (P_eos(rho) + p_shift) >= 0 on [rho_g, rho_l]  (:535-546)
    auto P_eos = [&](double r) { const double d = 1.0 - r; return r * R * prm.TT * (1.0 + (4.0 * r - 2.0 * r * r) / (d * d * d)) - a * r * r; };
    double worst = -1e30;
    for (int s = 0, Ns = 600; s <= Ns; ++s) {
        const double r = rhog + (rhol - rhog) * (double(s) / Ns);
        worst = std::max(worst, -((1.0 / 3.0) * r - P_eos(r)));
    }
    prm.p_shift = std::max(0.0, worst) + 1e-12;
    std::cout << "p_shift = " << std::setprecision(12) << prm.p_shift << "\n";
    auto psi = [&](double r) { const double S = (1.0 / 3.0) * r - (P_eos(r) + prm.p_shift); return S <= 0.0 ? 0.0 : std::sqrt(2.0 * S / (std::abs(G) * (1.0 / 3.0))); };
    std::cout << "psi(rho_l)=" << psi(rhol) << " psi(rho_g)=" << psi(rhog) << " psi(rho_w)=" << psi(rho_w) << "\n";

    DeviceLattice lat(prm);
    lat.init_case(CLBM_CASE_SC_LAYERED2D, {rhol, rhog, h_lower, (double)w_int});
    Stopwatch sw;
    std::ofstream energyfile("energy.dat"), mass_log("mass.dat");
    double M0 = -1.0;
    run_loop(lat, static_cast<int>(max_t / dt), out_freq, vtk_freq, sw, [&](int time_iter, bool vtk, bool out) {
        if (vtk) {
            auto f = lat.fields(true, true);
            const size_t n = lat.nelem();
            std::vector<double> fx(n), fy(n);
            check(clbm_download_force(lat.ctx, fx.data(), fy.data(), nullptr));
            VtkWriter w(time_iter, nx, ny, 1, dx);
            w.scalars("Density", "float", [&](size_t i) { return f.flag[i] == 0 ? 0.0 : f.s0[i]; });
            w.scalars("Pressure", "float", [&](size_t i) { return f.flag[i] == 0 ? 0.0 : f.s1[i]; });
            w.vectors("Force", [&](size_t i) { return std::array<double, 3>{fx[i], fy[i], 0.0}; });
            w.vectors("Velocity", [&](size_t i) { return std::array<double, 3>{f.flag[i] == 0 ? 0.0 : f.ux[i], f.flag[i] == 0 ? 0.0 : f.uy[i], 0.0}; });
        }
        if (!out) return;
        progress_line(time_iter, dt, max_t);
        const double M = lat.reduce(CLBM_REDUCE_MASS);
        if (M0 < 0.0) M0 = M;
        std::cout << std::setprecision(12) << "[Mass] M=" << M << "   \xCE\x94M/M0=" << std::setprecision(6) << (M - M0) / M0 * 100.0 << "%\n";
        if (mass_log) mass_log << std::setprecision(16) << time_iter * dt << " " << M << "\n";
        const double energy = lat.reduce(CLBM_REDUCE_ENERGY) * dx * dx / (dt * dt);
        std::cout << "Average energy: " << std::setprecision(10) << energy << "\n";
        energyfile << std::setw(10) << time_iter * dt << std::setw(16) << std::setprecision(10) << energy << "\n";
    });
    sw.report(lat.nelem());
}

}  // namespace coolbm
