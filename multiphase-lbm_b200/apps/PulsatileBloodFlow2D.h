// PulsatileBloodFlow2D.h -- compliant-vessel case on the B200 library.
// Driver surface of AB/apps/PulsatileBloodFlow2D.h:719-795 (PulsatileBloodFlow2D()): the parameters are hard-coded in
// the reference driver (N = 64, tau = 0.75, alpha = 0.01, severed, deformable); they are the defaults here and can be
// overridden on the command line.  tf = t_beat + 2 t_propagation iterations, a VTK file every tf/100 iterations with
// P, Ux, Uy cast to float and the Flag field (:680-706) -- byte-identical to the reference's files.
#pragma once
#include "case_common.h"

namespace coolbm {

inline void save_vtk_pulsatile(int nx, int ny, int time_iter, const std::vector<double> &P, const std::vector<double> &Ux,
                               const std::vector<double> &Uy, const std::vector<uint8_t> &flag)
{
    VtkWriter w(time_iter, nx, ny, 1, 1.0 / nx);
    w.scalars("P", "float", [&](size_t i) { return float(P[i]); });
    w.scalars("Ux", "float", [&](size_t i) { return float(Ux[i]); });
    w.scalars("Uy", "float", [&](size_t i) { return float(Uy[i]); });
    w.scalars("Flag", "int", [&](size_t i) { return flag[i] == 0 ? 1 : 0; });
}

inline void PulsatileBloodFlow2D(int N = 64, double tau = 0.75, double alpha = 0.01, bool is_severed = true, bool deformable = true,
                                 int max_iter = -1, bool vtk = true)
{
    clbm_pulsatile_params p{};
    p.abi_version = CLBM_ABI_VERSION;
    p.N = N; p.device = -1; p.is_severed = is_severed; p.deformable = deformable; p.t_beat = 0;
    p.tau = tau; p.alpha = alpha; p.p0_in = 0.20; p.p0_out = 0.19;
    clbm_pulsatile *sim = nullptr;
    if (clbm_pulsatile_create(&p, &sim) != CLBM_OK) throw std::runtime_error(clbm_last_error());   // "Initial wall location out of bounds."
    int nx, ny, tf;
    check(clbm_pulsatile_info(sim, &nx, &ny, &tf, nullptr, nullptr));
    if (max_iter >= 0) tf = std::min(tf, max_iter);
    const int step = std::max(1, tf / 100);
    const size_t nelem = (size_t)nx * ny;
    std::vector<double> P(nelem), Ux(nelem), Uy(nelem);
    std::vector<uint8_t> flag(nelem);
    Stopwatch sw;
    for (int t = 0; t <= tf;) {
        // the reference writes inside iteration t, after the wall update: that is the state after t + 1 iterations
        const int next = (t % step == 0) ? t : std::min(tf, (t / step + 1) * step);
        check(clbm_pulsatile_step(sim, next + 1 - sw.iters));
        sw.iters = next + 1;
        if (next % step == 0) {
            if (vtk) {
                check(clbm_pulsatile_download_fields(sim, P.data(), Ux.data(), Uy.data(), flag.data(), nullptr, nullptr));
                save_vtk_pulsatile(nx, ny, next, P, Ux, Uy, flag);
            }
            std::cout << "t=" << next << " / " << tf << "\n";
        }
        t = next + 1;
    }
    check(clbm_pulsatile_sync(sim));
    sw.report(nelem, "Runtime: ", "Throughput: ", " s\n");
    clbm_pulsatile_destroy(sim);
}

}  // namespace coolbm
