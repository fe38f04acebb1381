// rayleighTaylor2D.h -- HCZ phase-field Rayleigh-Taylor instability (N x 4N+2, walls y = 0, ny-1) on the B200 library.
// Driver surface of PF/apps/rayleighTaylor2D.h:862-986 (rayleighTaylor2D()): config keys, energy.dat,
// spike_bubble_position.dat (findInterfaceHeights :668-708, names swapped as in the reference), sol_*.vtk (:713-781).
#pragma once
#include "case_common.h"

namespace coolbm {

inline void print_hcz_parameters(const char *title, int N, int nx, int ny, int nz, double Re, double omega, double ulb, double max_t, double nu)
{
    std::cout << title << "\n" << "N      = " << N << '\n' << "nx     = " << nx << '\n' << "ny     = " << ny << '\n';
    if (nz > 1) std::cout << "nz     = " << nz << '\n';
    std::cout << "Re     = " << Re << '\n' << "omega  = " << omega << '\n' << "tau    = " << 1. / omega << '\n'
              << "nu     = " << nu << '\n' << "ulb    = " << ulb << '\n' << "max_t  = " << max_t << '\n';
}

struct HczConfig {
    double Re = 0, ulb = 0, max_t = 0, phi_l = 0, phi_g = 0, rho_l = 0, rho_g = 0, a = 0, b = 0, kappa = 0, gravity = 0;
    int N = 0, out_freq = 0, vtk_freq = 0;
    // optional keys of this library, absent from the reference's files: `collision MRT` selects the moment-space operator
    // (include/clbm.h, CLBM_COLLISION_MRT; HCZ D2Q9 only), s_e / s_eps / s_q its free rates (default: omega, i.e. BGK)
    bool mrt = false;
    double s_e = -1, s_eps = -1, s_q = -1;
    explicit HczConfig(Config cfg)
    {
        if (cfg.has("collision")) { cfg.used["collision"] = true; mrt = cfg.kv["collision"] == "MRT" || cfg.kv["collision"] == "mrt" || cfg.kv["collision"] == "1"; }
        s_e = cfg.d("s_e", -1); s_eps = cfg.d("s_eps", -1); s_q = cfg.d("s_q", -1);
        Re = cfg.d("Re", 0); ulb = cfg.d("ulb", 0); N = cfg.i("N", 0); max_t = cfg.d("max_t", 0); out_freq = cfg.i("out_freq", 0);
        vtk_freq = cfg.i("vtk_freq", 0); phi_l = cfg.d("phi_l", 0); phi_g = cfg.d("phi_g", 0); rho_l = cfg.d("rho_l", 0);
        rho_g = cfg.d("rho_g", 0); a = cfg.d("a", 0); b = cfg.d("b", 0); kappa = cfg.d("kappa", 0); gravity = cfg.d("gravity", 0);
        cfg.warn_unknown();
    }
    void fill(clbm_params &p, double omega) const
    {
        p.omega = omega; p.gravity = gravity; p.phi_l = phi_l; p.phi_g = phi_g; p.rho_l = rho_l; p.rho_g = rho_g;
        p.a = a; p.b = b; p.kappa = kappa;
        if (mrt && p.model == CLBM_MODEL_HCZ_D2Q9) {
            p.collision = CLBM_COLLISION_MRT;
            p.s_e = s_e > 0 ? s_e : omega; p.s_eps = s_eps > 0 ? s_eps : omega; p.s_q = s_q > 0 ? s_q : omega;
            std::cout << "collision = MRT  s_e = " << p.s_e << "  s_eps = " << p.s_eps << "  s_q = " << p.s_q << "  (s_nu = omega)\n";
        }
    }
};

// PF/apps/rayleighTaylor2D.h:668-708, scanned on the device (clbm_diag_interface_heights): no field download for two integers.
// The reference stores the x = 0 scan in `bubble_y` and the x = nx/2 scan in `spike_y`; kept.
inline void find_interface_heights(DeviceLattice &lat, double phi_l, double phi_g, int &spike_y, int &bubble_y)
{
    check(clbm_diag_interface_heights(lat.ctx, 0.5 * (phi_l + phi_g), &bubble_y, &spike_y));
}

inline void rayleighTaylor2D(const std::string &config_dir)
{
    HczConfig c(Config{read_config(config_dir + "/config_rayleighTaylor2D.txt", "config_rayleighTaylor2D.txt")});
    const int nx = c.N, ny = 4 * c.N + 2;
    const auto lb = lb_parameters(c.ulb, c.N, c.Re);
    print_hcz_parameters("Rayleigh-Taylor 2D problem", c.N, nx, ny, 1, c.Re, lb.omega, c.ulb, c.max_t, lb.nu);
    clbm_params prm = default_params(CLBM_MODEL_HCZ_D2Q9, nx, ny, 1);
    c.fill(prm, lb.omega);
    DeviceLattice lat(prm);
    lat.init_case(CLBM_CASE_HCZ_RT2D, {});

    Stopwatch sw;
    std::ofstream efile("energy.dat"), posfile("spike_bubble_position.dat"), velfile("spike_bubble_velocity.dat");
    const double dx = lb.dx, dt = lb.dt;
    run_loop(lat, static_cast<int>(c.max_t / dt), c.out_freq, c.vtk_freq, sw, [&](int time_iter, bool vtk, bool out) {
        if (vtk) {
            const DeviceLattice::Fields f = lat.fields(false, false);
            VtkWriter w(time_iter, nx, ny, 1, 1.0 / nx);   // the reference calls the writer without dx
            w.scalars("phi", "float", [&](size_t i) { return f.s0[i]; });
            w.scalars("density", "float", [&](size_t i) { return f.s2[i]; });
            w.scalars("Flag", "int", [&](size_t i) { return f.flag[i] == 0 ? "1" : "0"; }, true);
        }
        if (!out) return;
        progress_line(time_iter, dt, c.max_t, true);
        const double energy = lat.reduce(CLBM_REDUCE_ENERGY) * dx * dx / (dt * dt);
        std::cout << "Average energy: " << std::setprecision(8) << energy << std::endl;
        efile << std::setw(10) << time_iter * dt << std::setw(16) << std::setprecision(8) << energy << std::endl;
        int spike_y, bubble_y;
        find_interface_heights(lat, c.phi_l, c.phi_g, spike_y, bubble_y);
        posfile << std::setw(10) << time_iter * dt << std::setw(16) << (spike_y < 0 ? -1.0 : spike_y * dx) << std::setw(16)
                << (bubble_y < 0 ? -1.0 : bubble_y * dx) << "\n";
    });
    sw.report(lat.nelem());
}

}  // namespace coolbm
