// laplace3D.h -- HCZ phase-field droplet in a periodic N^3 box on the B200 library.
// Driver surface of PF/apps/laplace3D.h:852-950 (laplace3D()): config keys, energy.dat, sol_*.vtk with phi, Pressure
// (total_P) and Flag, z planes written from nz-1 down (:689-783), "Runtime / MLUPS" report (:101-112).
#pragma once
#include "rayleighTaylor2D.h"

namespace coolbm {

inline void laplace3D(const std::string &config_dir)
{
    HczConfig c(Config{read_config(config_dir + "/config_laplace3D.txt", "config_laplace3D.txt")});
    const int N = c.N;
    const auto lb = lb_parameters(c.ulb, N, c.Re);
    print_hcz_parameters("laplace 3-D problem", N, N, N, N, c.Re, lb.omega, c.ulb, c.max_t, lb.nu);
    clbm_params prm = default_params(CLBM_MODEL_HCZ_D3Q19, N, N, N);
    c.fill(prm, lb.omega);
    DeviceLattice lat(prm);
    lat.init_case(CLBM_CASE_HCZ_LAPLACE3D, {});

    Stopwatch sw;
    std::ofstream efile("energy.dat");
    const double dx = lb.dx, dt = lb.dt;
    run_loop(lat, static_cast<int>(c.max_t / dt), c.out_freq, c.vtk_freq, sw, [&](int time_iter, bool vtk, bool out) {
        if (vtk) {
            auto f = lat.fields(true, false);
            VtkWriter w(time_iter, N, N, N, 1.0 / N);
            w.scalars("phi", "float", [&](size_t i) { return f.s0[i]; });
            w.scalars("Pressure", "float", [&](size_t i) { return f.s1[i]; });
            w.scalars("Flag", "int", [&](size_t i) { return f.flag[i] == 0 ? "1" : "0"; }, true);
        }
        if (!out) return;
        progress_line(time_iter, dt, c.max_t, true);
        const double energy = lat.reduce(CLBM_REDUCE_ENERGY) * dx * dx / (dt * dt);
        std::cout << "Average energy: " << std::setprecision(8) << energy << std::endl;
        efile << std::setw(10) << time_iter * dt << std::setw(16) << std::setprecision(8) << energy << std::endl;
    });
    sw.report(lat.nelem(), "Runtime: ", "Throughput: ");
}

}  // namespace coolbm
