// Harness around SC/apps/RayleighTaylor2D.h (untouched; commented out of the shipped COOLBM.cpp :74, so it is compiled
// here on its own).  Setup mirrors RayleighTaylor2D() :577-640 with the lattice extent and omega passed in.
// Dumped: "in" populations, density (:172-183), P_eos (:200-208), u_eq = u + F/(2 rho) (:343-351, what
// computeEnergy_RayleighTaylor2D :503-516 sums), force_ff (:236-289) at bulk nodes, flag.
#include "harness_common.h"
#include <cstdint>
#include "RayleighTaylor2D.h"
int main(int argc, char** argv)
{
    Args A(argc, argv);
    int nx = A.i("nx", 16), ny = A.i("ny", 66), steps = A.i("steps", 10), threads = A.i("threads", 1);
    double omega = A.d("omega", 1.0), rhol = A.d("rhol", 1.2), rhog = A.d("rhog", 0.4), rhow = A.d("rhow", 0.2);
    double g = A.d("g", -5.0), a = A.d("a", 1.0), b = A.d("b", 4.0), gravity = A.d("gravity", -1.25e-5);
    Dim_RayleighTaylor2D dim{nx, ny};
    vector<double> lattice_vect(LBM_RayleighTaylor2D::sizeOfLattice(dim.nelem));
    double* lattice = lattice_vect.data();
    vector<CellType_RayleighTaylor2D> flag_vect(dim.nelem);
    vector<int> parity_vect{0};
    int* parity = &parity_vect[0];
    auto [c, opp, t] = d2q9_constants_RayleighTaylor2D();
    LBM_RayleighTaylor2D lbm{lattice, flag_vect.data(), parity, &c[0], &opp[0], &t[0], omega, rhol, rhog, rhow, g, a, b, gravity, dim};
    for_each(lattice, lattice + dim.nelem, [&lbm](double& f0) { lbm.iniLattice(f0); });
    inigeom_RayleighTaylor2D(lbm);
    double sec = run_steps(lbm, lattice, dim.nelem, parity, steps, threads);
    report("sc_rt2d", dim.nelem, steps, threads, sec);
    std::printf("{\"energy\": %.17g}\n", computeEnergy_RayleighTaylor2D(lbm));
    Dump D(A.s("out", ""));
    if (D.f) {
        D.put(lattice + (size_t)(*parity) * dim.npop, dim.npop);
        vector<double> rho(dim.nelem), pr(dim.nelem, 0.0), ux(dim.nelem, 0.0), uy(dim.nelem, 0.0), fx(dim.nelem, 0.0), fy(dim.nelem, 0.0);
        for (size_t i = 0; i < dim.nelem; ++i) {
            rho[i] = lbm.density(lattice[i]);
            if (flag_vect[i] != CellType_RayleighTaylor2D::bulk) continue;
            pr[i] = lbm.P_eos(lattice[i]);
            auto u = lbm.u_eq(lattice[i]);
            ux[i] = u[0]; uy[i] = u[1];
            auto F = lbm.force_ff(lattice[i]);
            fx[i] = F[0]; fy[i] = F[1];
        }
        D.put(rho); D.put(pr); D.put(ux); D.put(uy); D.put(fx); D.put(fy);
        D.put_u8((uint8_t*)flag_vect.data(), dim.nelem);
    }
    return 0;
}
