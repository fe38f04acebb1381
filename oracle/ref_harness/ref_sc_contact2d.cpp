// Harness around SC/apps/contactAngle2D.h (untouched).  Setup mirrors contactAngle2D() :708-763.
#include "harness_common.h"
#include "contactAngle2D.h"
int main(int argc, char** argv)
{
    Args A(argc, argv);
    int nx = A.i("nx", 48), ny = A.i("ny", 24), steps = A.i("steps", 10), threads = A.i("threads", 1);
    double omega = A.d("omega", 1.0), rhol = A.d("rhol", 0.265), rhog = A.d("rhog", 0.038), rho_w = A.d("rho_w", 0.2);
    double a = A.d("a", 1.0), b = A.d("b", 4.0), R = A.d("R", 1.0), TT0 = A.d("TT0", 0.875), gravity = A.d("gravity", 0.0);
    double RR = A.d("RR", 8.0);
    Dim_contactAngle2D dim{nx, ny};
    vector<double> lattice_vect(LBM_contactAngle2D::sizeOfLattice(dim.nelem));
    double* lattice = lattice_vect.data();
    vector<CellType_contactAngle2D> flag_vect(dim.nelem);
    vector<int> parity_vect{0};
    int* parity = &parity_vect[0];
    auto [c, opp, t] = d2q9_constants_contactAngle2D();
    LBM_contactAngle2D lbm{lattice, flag_vect.data(), parity, &c[0], &opp[0], &t[0], omega, rhol, rhog, rho_w, a, b, R, TT0, 0.0, gravity, RR, dim};
    lbm.TT = lbm.TT0 * (0.3773 * a / (b * R));
    for_each(lattice, lattice + dim.nelem, [&lbm](double& f0) { lbm.iniLattice(f0); });
    inigeom_contactAngle2D(lbm);
    double sec = run_steps(lbm, lattice, dim.nelem, parity, steps, threads);
    report("sc_contact2d", dim.nelem, steps, threads, sec);
    Dump D(A.s("out", ""));
    if (D.f) {
        D.put(lattice + (size_t)(*parity) * dim.npop, dim.npop);
        vector<double> rho(dim.nelem), pr(dim.nelem, 0.0), ux(dim.nelem, 0.0), uy(dim.nelem, 0.0);
        for (size_t i = 0; i < dim.nelem; ++i) {
            rho[i] = lbm.density(lattice[i]);
            if (flag_vect[i] != CellType_contactAngle2D::bulk) continue;
            pr[i] = lbm.pressure_node(lattice[i]);
            auto u = lbm.u_actual(lattice[i]);
            ux[i] = u[0]; uy[i] = u[1];
        }
        D.put(rho); D.put(pr); D.put(ux); D.put(uy);
        D.put_u8((uint8_t*)flag_vect.data(), dim.nelem);
    }
    return 0;
}
