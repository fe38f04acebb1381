// Harness around SC/apps/twoLayeredFlow2D.h (untouched).  Setup mirrors twoLayeredFlow2D() :492-556, with the lattice
// extent, omega and p_shift passed in (p_shift = the driver's 601-point scan, computed by the caller).
#include "harness_common.h"
#include "twoLayeredFlow2D.h"
int main(int argc, char** argv)
{
    Args A(argc, argv);
    int nx = A.i("nx", 10), ny = A.i("ny", 41), steps = A.i("steps", 10), threads = A.i("threads", 1), w_int = A.i("w_int", 4);
    double omega = A.d("omega", 1.0), rhol = A.d("rhol", 0.21), rhog = A.d("rhog", 0.067), rho_w = A.d("rho_w", 0.067);
    double a = A.d("a", 1.0), b = A.d("b", 4.0), R = A.d("R", 1.0), TT0 = A.d("TT0", 0.95);
    double gx = A.d("gx", 1e-8), gy = A.d("gy", 0.0), G = A.d("G", -1.0), h_lower = A.d("h_lower", 0.3);
    Dim dim{nx, ny};
    vector<double> lattice_vect(LBM::sizeOfLattice(dim.nelem));
    double* lattice = lattice_vect.data();
    vector<CellType> flag_vect(dim.nelem);
    vector<int> parity_vect{0};
    int* parity = &parity_vect[0];
    auto [c, opp, t] = d2q9_constants();
    LBM lbm{lattice, flag_vect.data(), parity, &c[0], &opp[0], &t[0], omega, rhol, rhog, rho_w, a, b, R, TT0, 0.0, gx, gy, G, 0.0, dim};
    lbm.TT = lbm.TT0 * (0.3773 * a / (b * R));
    {   // p_shift exactly as the driver chooses it (:535-546)
        double worst = -1e30; int Ns = 600;
        for (int s = 0; s <= Ns; ++s) {
            double r = rhog + (rhol - rhog) * (double(s) / Ns);
            double S = lbm.cs2() * r - lbm.P_eos_rho(r);
            worst = std::max(worst, -S);
        }
        lbm.p_shift = std::max(0.0, worst) + 1e-12;
    }
    for_each(lattice, lattice + dim.nelem, [&lbm, h_lower, w_int](double& f0) { lbm.iniLattice_layers(f0, h_lower, w_int); });
    inigeom(lbm);
    double sec = run_steps(lbm, lattice, dim.nelem, parity, steps, threads);
    report("sc_layered2d", dim.nelem, steps, threads, sec);
    std::printf("{\"p_shift\": %.17g}\n", lbm.p_shift);
    Dump D(A.s("out", ""));
    if (D.f) {
        D.put(lattice + (size_t)(*parity) * dim.npop, dim.npop);
        vector<double> rho(dim.nelem), pr(dim.nelem, 0.0), ux(dim.nelem, 0.0), uy(dim.nelem, 0.0);
        for (size_t i = 0; i < dim.nelem; ++i) {
            rho[i] = lbm.density(lattice[i]);
            if (flag_vect[i] != CellType::bulk) continue;
            pr[i] = lbm.pressure_node(lattice[i]);
            auto u = lbm.u_actual(lattice[i]);
            ux[i] = u[0]; uy[i] = u[1];
        }
        D.put(rho); D.put(pr); D.put(ux); D.put(uy);
        D.put_u8((uint8_t*)flag_vect.data(), dim.nelem);
    }
    return 0;
}
