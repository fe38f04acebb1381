// contactAngle2D.h -- Shan-Chen droplet on a wetting wall (2N x N, walls y = 0, ny-1) on the B200 library.
// Driver surface of SC/apps/contactAngle2D.h:644-806 (contactAngle2D()) incl. the base/height contact-angle
// measurement (:465-529), whose scans run on the device.
#pragma once
#include <cmath>

#include "laplace2D.h"

namespace coolbm {

// base/height method, threshold 0.5 (rho_l + rho_g): the three scans run on the device (clbm_diag_contact_angle), the
// circle geometry and the log lines are those of calculateContactAngle (:507-528)
inline void calculate_contact_angle(DeviceLattice &lat, int ny, double rho_l, double rho_g)
{
    const double PI = 3.14159265358979323846;
    int base_y = 0, base = 0, height = 0;
    check(clbm_diag_contact_angle(lat.ctx, 0.5 * (rho_l + rho_g), &base_y, &base, &height));
    if (base_y >= ny - 1) { std::cout << "ContactAngle: no fluid row found above wall.\n"; return; }
    if (height <= 0 || base <= 1) {
        std::cout << "ContactAngle: droplet not detected (Base=" << base << ", Height=" << height << ")\n";
        return;
    }
    const double h = height, b = base, R = (4.0 * h * h + b * b) / (8.0 * h);
    double theta = std::atan((0.5 * b) / (R - h)) * 180.0 / PI;
    if (theta < 0.0) theta += 180.0;
    std::cout << std::setprecision(8) << "Base=" << base << " Height=" << height << " ContactAngle=" << theta << " deg\n";
    static std::ofstream ca_log("contact_angle.dat", std::ios::app);
    if (ca_log) ca_log << base << " " << height << " " << theta << "\n";
}

inline void contactAngle2D(const std::string &config_dir)
{
    Config cfg{read_config(config_dir + "/config_contactAngle2D.txt", "config_contactAngle2D.txt")};
    const double Re = cfg.d("Re", 60), ulb = cfg.d("ulb", 0.1), max_t = cfg.d("max_t", 10.0), rhol = cfg.d("rhol", 1.0),
                 rhog = cfg.d("rhog", 0.1), a = cfg.d("a", 1.0), b = cfg.d("b", 4.0), R = cfg.d("R", 1.0), TT0 = cfg.d("TT0", 0.875),
                 gravity = cfg.d("gravity", 0.0), tau_in = cfg.d("tau", -1.0), RR = cfg.d("RR", 100);
    const double rho_w = cfg.has("rhow") ? cfg.d("rhow", 0.12) : cfg.d("rho_w", 0.12);
    cfg.i("k_index", -1);
    const int N = cfg.i("N", 100), out_freq = cfg.i("out_freq", 400), vtk_freq = cfg.i("vtk_freq", 400);
    const int nx = 2 * N, ny = N;

    double nu, omega, dx = 1.0 / N, dt = dx * ulb;
    if (tau_in > 0.0) { omega = 1.0 / tau_in; nu = (tau_in - 0.5) / 3.0; }
    else { auto p = lb_parameters(ulb, N, Re); nu = p.nu; omega = p.omega; dx = p.dx; dt = p.dt; }
    print_sc_parameters("Contact angle 2D problem", N, nx, ny, Re, omega, ulb, max_t, nu);

    clbm_params prm = default_params(CLBM_MODEL_SC_D2Q9, nx, ny, 1);
    const double Tc = 0.3773 * a / (b * R);
    prm.omega = omega; prm.gravity = gravity; prm.rho_w = rho_w; prm.a = a; prm.b = b; prm.R = R; prm.TT = TT0 * Tc;
    prm.sc_force = CLBM_SC_FORCE_CONTACT;
    apply_collision_keys(cfg, prm, omega);
    std::cout << std::setprecision(6) << "CS params: a=" << a << " b=" << b << " R=" << R << "\n"
              << "TT0 (reduced)=" << TT0 << "  Tc=" << Tc << "  TT (abs)=" << prm.TT << "\n"
              << "rho_l=" << rhol << "  rho_g=" << rhog << "  rho_w=" << rho_w << "  RR=" << RR << "\n"
              << "gravity=" << gravity << "\n";
    DeviceLattice lat(prm);
    lat.init_case(CLBM_CASE_SC_CONTACT2D, {rhol, rhog, RR});

    Stopwatch sw;
    std::ofstream energyfile("energy.dat"), mass_log("mass.dat");
    double M0 = -1.0;
    const int max_time_iter = static_cast<int>(max_t / dt);
    run_loop(lat, max_time_iter, out_freq, vtk_freq, sw, [&](int time_iter, bool vtk, bool out) {
        if (vtk) save_vtk_sc(lat, time_iter, dx);
        if (!out) return;
        progress_line(time_iter, dt, max_t);
        calculate_contact_angle(lat, ny, rhol, rhog);
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

// BASELINE configs[3]: the same physics on D3Q19, sessile droplet between bounce-back planes y = 0, ny-1 (N^3).
// The reference ships no D3Q19 Shan-Chen driver; this one follows contactAngle2D() key for key.
inline void droplet3D(const std::string &config_dir)
{
    Config cfg{read_config(config_dir + "/config_droplet3D.txt", "config_droplet3D.txt")};
    const double ulb = cfg.d("ulb", 0.1), max_t = cfg.d("max_t", 1.0), rhol = cfg.d("rhol", 0.265), rhog = cfg.d("rhog", 0.038),
                 a = cfg.d("a", 1.0), b = cfg.d("b", 4.0), R = cfg.d("R", 1.0), TT0 = cfg.d("TT0", 0.875), tau_in = cfg.d("tau", 1.0),
                 Re = cfg.d("Re", 60);
    const double rho_w = cfg.has("rhow") ? cfg.d("rhow", 0.2) : cfg.d("rho_w", 0.2);
    const int N = cfg.i("N", 128), out_freq = cfg.i("out_freq", 100), vtk_freq = cfg.i("vtk_freq", 0);
    const double RR = cfg.d("RR", 0.2 * N), yc = cfg.d("yc", 5.0);
    double nu, omega, dx = 1.0 / N, dt = dx * ulb;
    if (tau_in > 0.0) { omega = 1.0 / tau_in; nu = (tau_in - 0.5) / 3.0; }
    else { auto p = lb_parameters(ulb, N, Re); nu = p.nu; omega = p.omega; dx = p.dx; dt = p.dt; }
    print_sc_parameters("Sessile droplet 3D problem (D3Q19 Shan-Chen)", N, N, N, Re, omega, ulb, max_t, nu);
    clbm_params prm = default_params(CLBM_MODEL_SC_D3Q19, N, N, N);
    prm.omega = omega; prm.rho_w = rho_w; prm.a = a; prm.b = b; prm.R = R; prm.TT = TT0 * (0.3773 * a / (b * R));
    prm.sc_force = CLBM_SC_FORCE_CONTACT;
    DeviceLattice lat(prm);
    lat.init_case(CLBM_CASE_SC_DROPLET3D, {rhol, rhog, RR, yc});
    Stopwatch sw;
    std::ofstream energyfile("energy.dat"), mass_log("mass.dat");
    double M0 = -1.0;
    run_loop(lat, static_cast<int>(max_t / dt), out_freq, vtk_freq, sw, [&](int time_iter, bool vtk, bool out) {
        if (vtk) save_vtk_sc(lat, time_iter, dx);
        if (!out) return;
        progress_line(time_iter, dt, max_t);
        const double M = lat.reduce(CLBM_REDUCE_MASS);
        if (M0 < 0.0) M0 = M;
        std::cout << std::setprecision(12) << "[Mass] M=" << M << "   \xCE\x94M/M0=" << std::setprecision(6) << (M - M0) / M0 * 100.0 << "%\n";
        if (mass_log) mass_log << std::setprecision(16) << time_iter * dt << " " << M << "\n";
        const double energy = lat.reduce(CLBM_REDUCE_ENERGY) * dx * dx / (dt * dt);
        std::cout << "Average energy: " << std::setprecision(10) << energy << "  max|u|: " << lat.reduce(CLBM_REDUCE_UMAX) << "\n";
        energyfile << std::setw(10) << time_iter * dt << std::setw(16) << std::setprecision(10) << energy << "\n";
    });
    sw.report(lat.nelem());
}

}  // namespace coolbm
