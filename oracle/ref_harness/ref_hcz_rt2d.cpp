// Harness around PF/apps/rayleighTaylor2D.h (untouched).  Setup mirrors rayleighTaylor2D() :905-932.
#include "harness_common.h"
#include "rayleighTaylor2D.h"
int main(int argc, char** argv)
{
    Args A(argc, argv);
    int nx = A.i("nx", 16), ny = A.i("ny", 4 * nx + 2), steps = A.i("steps", 10), threads = A.i("threads", 1);
    double omega = A.d("omega", 1.0), phi_l = A.d("phi_l", 0.251), phi_g = A.d("phi_g", 0.024);
    double rho_l = A.d("rho_l", 0.12), rho_g = A.d("rho_g", 0.04), a = A.d("a", 4.0), b = A.d("b", 4.0);
    double kappa = A.d("kappa", 0.01), gravity = A.d("gravity", -6.25e-6);
    Dim_rayleighTaylor2D dim{nx, ny};
    vector<CellData> lattice_vect(LBM_rayleighTaylor2D::sizeOfLattice(dim.nelem));
    CellData* lattice = &lattice_vect[0];
    vector<CellType_rayleighTaylor2D> flag_vect(dim.nelem);
    vector<int> parity_vect{0};
    int* parity = &parity_vect[0];
    auto [c_vect, opp_vect, t_vect] = d2q9_constants_rayleighTaylor2D();
    LBM_rayleighTaylor2D lbm{lattice, &flag_vect[0], parity, &c_vect[0], &opp_vect[0], &t_vect[0],
                             omega, phi_l, phi_g, rho_l, rho_g, a, b, kappa, gravity, dim};
    for_each(lattice, lattice + dim.nelem, [&lbm](CellData& f0) { lbm.iniLattice(f0); });
    inigeom_rayleighTaylor2D(lbm);
    double sec = run_steps(lbm, lattice, dim.nelem, parity, steps, threads);
    report("hcz_rt2d", dim.nelem, steps, threads, sec);
    Dump D(A.s("out", ""));
    if (D.f) {
        D.put(lattice + (size_t)(*parity) * dim.npop, dim.npop);                 // f "in"
        D.put(lattice + 2 * dim.npop + (size_t)(*parity) * dim.npop, dim.npop);  // g "in"
        vector<double> phi(dim.nelem), P(dim.nelem, 0.0), rho(dim.nelem), ux(dim.nelem, 0.0), uy(dim.nelem, 0.0);
        for (size_t i = 0; i < dim.nelem; ++i) {
            auto [ph, pt] = lbm.macro_phi_P(lattice[i]);
            phi[i] = ph;
            rho[i] = lbm.total_rho(lattice[i]);
            if (flag_vect[i] != CellType_rayleighTaylor2D::bulk) continue;
            P[i] = lbm.total_P(lattice[i]);
            auto u = lbm.velocity(lattice[i]);
            ux[i] = u[0]; uy[i] = u[1];
        }
        D.put(phi); D.put(P); D.put(rho); D.put(ux); D.put(uy);
        D.put_u8((uint8_t*)&flag_vect[0], dim.nelem);
    }
    return 0;
}
