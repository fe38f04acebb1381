// Harness around SC/apps/laplace2D.h (untouched).  Setup mirrors Laplace2D() :439-474.
#include "harness_common.h"
#include "laplace2D.h"
int main(int argc, char** argv)
{
    Args A(argc, argv);
    int nx = A.i("nx", 32), ny = A.i("ny", nx), steps = A.i("steps", 10), threads = A.i("threads", 1);
    double omega = A.d("omega", 1.0), rhol = A.d("rhol", 0.265), rhog = A.d("rhog", 0.038), rho_w = A.d("rho_w", 0.12);
    double a = A.d("a", 1.0), b = A.d("b", 4.0), R = A.d("R", 1.0), TT0 = A.d("TT0", 0.875), gravity = A.d("gravity", 0.0);
    Dim_Laplace2D dim{nx, ny};
    vector<double> lattice_vect(LBM_Laplace2D::sizeOfLattice(dim.nelem));
    double* lattice = &lattice_vect[0];
    vector<CellType_Laplace2D> flag_vect(dim.nelem);
    vector<int> parity_vect{0};
    int* parity = &parity_vect[0];
    auto [c, opp, t] = d2q9_constants_Laplace2D();
    LBM_Laplace2D lbm{lattice, &flag_vect[0], parity, &c[0], &opp[0], &t[0], omega, rhol, rhog, rho_w, a, b, R, TT0, 0.0, gravity, dim};
    lbm.TT = lbm.TT0 * (0.3773 * a / (b * R));
    for_each(lattice, lattice + dim.nelem, [&lbm](double& f0) { lbm.iniLattice(f0); });
    inigeom_Laplace2D(lbm);
    double sec = run_steps(lbm, lattice, dim.nelem, parity, steps, threads);
    report("sc_laplace2d", dim.nelem, steps, threads, sec);
    Dump D(A.s("out", ""));
    if (D.f) {
        D.put(lattice + (size_t)(*parity) * dim.npop, dim.npop);
        vector<double> rho(dim.nelem), pr(dim.nelem), ux(dim.nelem), uy(dim.nelem);
        for (size_t i = 0; i < dim.nelem; ++i) {
            rho[i] = lbm.density(lattice[i]);
            pr[i] = lbm.pressure_node(lattice[i]);
            auto u = lbm.u_actual(lattice[i]);
            ux[i] = u[0]; uy[i] = u[1];
        }
        D.put(rho); D.put(pr); D.put(ux); D.put(uy);
        D.put_u8((uint8_t*)&flag_vect[0], dim.nelem);
    }
    return 0;
}
