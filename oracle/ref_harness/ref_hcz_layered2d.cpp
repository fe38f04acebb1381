// Harness around PF/apps/twoLayeredFlow2D.h (untouched).  Setup mirrors twoLayered2D() :805-823.
#include "harness_common.h"
#include "twoLayeredFlow2D.h"
int main(int argc, char** argv)
{
    Args A(argc, argv);
    int nx = A.i("nx", 10), ny = A.i("ny", 41), w_int = A.i("w_int", 2), steps = A.i("steps", 10), threads = A.i("threads", 1);
    double omega = A.d("omega", 1.0), phi_l = A.d("phi_l", 0.251), phi_g = A.d("phi_g", 0.024);
    double rho_l = A.d("rho_l", 0.12), rho_g = A.d("rho_g", 0.04), a = A.d("a", 4.0), b = A.d("b", 4.0);
    double kappa = A.d("kappa", 0.001), gx = A.d("gx", 0.0), Gx_const = A.d("gx_const", 1e-8), h_lower = A.d("h_lower", 0.3);
    Dim_twoLayeredPF2D dim{nx, ny};
    vector<CellData> lattice_vect(LBM_twoLayeredPF2D::sizeOfLattice(dim.nelem));
    CellData* lattice = &lattice_vect[0];
    vector<CellType_twoLayeredPF2D> flag_vect(dim.nelem);
    vector<int> parity_vect{0};
    int* parity = &parity_vect[0];
    auto [c_vect, opp_vect, t_vect] = d2q9_constants_twoLayeredPF2D();
    LBM_twoLayeredPF2D lbm{lattice, &flag_vect[0], parity, &c_vect[0], &opp_vect[0], &t_vect[0],
                             omega, phi_l, phi_g, rho_l, rho_g, a, b, kappa, gx, Gx_const, dim};
    for_each(lattice, lattice + dim.nelem, [&lbm, h_lower, w_int](CellData& f0) { lbm.iniLattice_layers(f0, h_lower, w_int); });
    inigeom_twoLayeredPF2D(lbm);
    double sec = run_steps(lbm, lattice, dim.nelem, parity, steps, threads);
    report("hcz_layered2d", dim.nelem, steps, threads, sec);
    Dump D(A.s("out", ""));
    if (D.f) {
        D.put(lattice + (size_t)(*parity) * dim.npop, dim.npop);                 // f "in"
        D.put(lattice + 2 * dim.npop + (size_t)(*parity) * dim.npop, dim.npop);  // g "in"
        vector<double> phi(dim.nelem), P(dim.nelem, 0.0), rho(dim.nelem), ux(dim.nelem, 0.0), uy(dim.nelem, 0.0);
        for (size_t i = 0; i < dim.nelem; ++i) {
            auto [ph, pt] = lbm.macro_phi_P(lattice[i]);
            phi[i] = ph;
            rho[i] = lbm.total_rho(lattice[i]);
            if (flag_vect[i] != CellType_twoLayeredPF2D::bulk) continue;
            P[i] = lbm.total_P(lattice[i]);
            auto u = lbm.velocity(lattice[i]);
            ux[i] = u[0]; uy[i] = u[1];
        }
        D.put(phi); D.put(P); D.put(rho); D.put(ux); D.put(uy);
        D.put_u8((uint8_t*)&flag_vect[0], dim.nelem);
    }
    return 0;
}
