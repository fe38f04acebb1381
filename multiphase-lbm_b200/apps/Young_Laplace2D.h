// Young_Laplace2D.h -- conservative phase-field (Fakhari) bubble in a periodic N x N box on the B200 library: the problem
// the reference's AB build runs by default (AB/apps/COOLBM.cpp:99).  Driver surface of AB/apps/Young_Laplace2D.h:456-568
// (Young_Laplace2D()): config_laplace2D.txt keys (N, tf, out_freq, vtk_freq, Sigma, W, M, RhoL, RhoH, tau), "it = ..."
// progress lines, energy.dat / mass.dat, sol_*.vtk with phi, Pressure, velocity, Flag cast to float (:374-421).
#pragma once
#include <array>

#include "twoLayeredFlow2D.h"

namespace coolbm {

inline void Young_Laplace2D(const std::string &config_dir)
{
    Config cfg{read_config_lines(config_dir + "/config_laplace2D.txt",
                                 "Config file not found. Expected \"config_laplace2D.txt\" in ../apps/Config_Files/")};
    const int N = cfg.i("N", 128);
    int tf = cfg.i("tf", 10000);
    const int out_freq = cfg.i("out_freq", tf / 10), vtk_freq = cfg.i("vtk_freq", tf / 10);
    clbm_yl2d_params p{};
    p.abi_version = CLBM_ABI_VERSION;
    p.nx = N; p.ny = N; p.device = -1;
    p.Sigma = cfg.d("Sigma", 0.01); p.W = cfg.d("W", 4.0); p.M = cfg.d("M", 0.02);
    p.RhoL = cfg.d("RhoL", 0.001); p.RhoH = cfg.d("RhoH", 1.0); p.tau = cfg.d("tau", 0.8);
    cfg.warn_unknown();
    clbm_yl2d *sim = nullptr;
    check(clbm_yl2d_create(&p, &sim));
    const size_t nelem = (size_t)N * N;
    std::vector<double> C(nelem), P(nelem), Ux(nelem), Uy(nelem);

    Stopwatch sw;
    std::ofstream efile("energy.dat"), mass_log("mass.dat");
    double M0 = -1.0;
    for (int it = 0; it <= tf;) {
        const bool vtk = vtk_freq != 0 && it % vtk_freq == 0, out = out_freq != 0 && it % out_freq == 0;
        if (vtk) {
            check(clbm_yl2d_download_fields(sim, C.data(), P.data(), nullptr, Ux.data(), Uy.data()));
            VtkWriter w(it, N, N, 1, 1.0);
            w.scalars("phi", "float", [&](size_t i) { return float(C[i]); });
            w.scalars("Pressure", "float", [&](size_t i) { return float(P[i]); });
            w.vectors_rows("velocity", [&](size_t i) { return std::array<float, 2>{float(Ux[i]), float(Uy[i])}; });
            w.scalars("Flag", "int", [&](size_t) { return 0; });
        }
        if (out) {
            std::cout << "it = " << std::setw(7) << it << "   [" << std::fixed << std::setprecision(1) << (100.0 * it / double(tf)) << "%]" << std::endl;
            double energy = 0, Mcur = 0;
            check(clbm_yl2d_reduce(sim, CLBM_REDUCE_ENERGY, &energy));
            check(clbm_yl2d_reduce(sim, CLBM_REDUCE_MASS, &Mcur));
            std::cout << "Average kinetic energy: " << std::setprecision(8) << energy << std::endl;
            if (efile) efile << std::setw(10) << it << std::setw(16) << std::setprecision(8) << energy << "\n";
            if (M0 < 0.0) M0 = Mcur;
            std::cout << std::setprecision(12) << "[Mass] M=" << Mcur << "   \xCE\x94M/M0=" << std::setprecision(6) << (Mcur - M0) / M0 * 100.0 << "%\n";
            if (mass_log) mass_log << it << " " << std::setprecision(16) << Mcur << "\n";
        }
        int next = tf + 1;                                  // the loop runs it = 0 .. tf inclusive (:539)
        if (vtk_freq != 0) next = std::min(next, (it / vtk_freq + 1) * vtk_freq);
        if (out_freq != 0) next = std::min(next, (it / out_freq + 1) * out_freq);
        check(clbm_yl2d_step(sim, next - it));
        sw.iters += next - it;
        it = next;
    }
    check(clbm_yl2d_sync(sim));
    sw.report(nelem, "Runtime: ", "Throughput: ", " s\n");
    clbm_yl2d_destroy(sim);
}

}  // namespace coolbm
