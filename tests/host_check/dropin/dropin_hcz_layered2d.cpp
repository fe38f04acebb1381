// Drop-in test around PF/apps/twoLayeredFlow2D.h (untouched).  Its iniLattice_layers fills BOTH lattice buffers (:184-187);
// clbm_upload hands the bounce_back-node values of the buffer the parity does not select over too, so after an ODD number
// of steps clbm_download_lattice must return the reference's host array at EVERY node, walls included.
#include "dropin_common.h"
#include "twoLayeredFlow2D.h"
int main(int argc, char** argv)
{
    Args A(argc, argv);
    int nx = A.i("nx", 10), ny = A.i("ny", 41), w_int = A.i("w_int", 2), steps = A.i("steps", 301), threads = A.i("threads", 1);
    double omega = A.d("omega", 1.0), phi_l = A.d("phi_l", 0.251), phi_g = A.d("phi_g", 0.024);
    double rho_l = A.d("rho_l", 0.12), rho_g = A.d("rho_g", 0.04), a = A.d("a", 4.0), b = A.d("b", 4.0);
    double kappa = A.d("kappa", 0.001), gx = A.d("gx", 1e-7), Gx_const = A.d("gx_const", 1e-6), h_lower = A.d("h_lower", 0.3);
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

    vector<CellData> lattice2_vect(lattice_vect);
    vector<int> parity2_vect{*parity};
    LBM_twoLayeredPF2D lbm2 = lbm;
    lbm2.lattice = &lattice2_vect[0];
    lbm2.parity = &parity2_vect[0];

    clbm_params p{};                         // scalar members of LBM_twoLayeredPF2D (twoLayeredFlow2D.h:112-130)
    p.abi_version = CLBM_ABI_VERSION;  p.model = CLBM_MODEL_HCZ_D2Q9;  p.sc_force = CLBM_HCZ_FORCE_LAYERED;
    p.nx = p.nx_global = dim.nx;  p.ny = dim.ny;  p.nz = 1;  p.x_offset = 0;  p.device = -1;  p.fused = 1;
    p.omega = lbm.omega;
    p.phi_l = lbm.phi_l;  p.phi_g = lbm.phi_g;  p.rho_l = lbm.rho_l;  p.rho_g = lbm.rho_g;  p.a = lbm.a;  p.b = lbm.b;  p.kappa = lbm.kappa;
    p.gx = lbm.gx;  p.gx_const = lbm.Gx_const;
    clbm_ctx* ctx = nullptr;
    DROPIN_CLBM(clbm_create(&p, &ctx));
    DROPIN_CLBM(clbm_upload(ctx, lbm2.lattice, reinterpret_cast<const uint8_t*>(&flag_vect[0]), *lbm2.parity));
    DROPIN_CLBM(clbm_step(ctx, steps));
    DROPIN_CLBM(clbm_download_lattice(ctx, lbm2.lattice, lbm2.parity));
    DROPIN_CLBM(clbm_destroy(ctx));

    run_steps(lbm, lattice, dim.nelem, parity, steps, threads);

    ErrList E(A.d("tol", 1e-10));
    E.exact("parity", *lbm2.parity == *parity);
    const size_t n = dim.nelem;
    vector<double> phi1(n), phi2(n), rho1(n), rho2(n), ux1(n, 0.), ux2(n, 0.), uy1(n, 0.), uy2(n, 0.);
    bool walls_equal = true;
    size_t nwall = 0;
    for (size_t i = 0; i < n; ++i) {
        auto [ph1, pt1] = lbm.macro_phi_P(lattice[i]);
        auto [ph2, pt2] = lbm2.macro_phi_P(lbm2.lattice[i]);
        phi1[i] = ph1; phi2[i] = ph2;                                   // ALL nodes, bounce_back ones included
        rho1[i] = lbm.total_rho(lattice[i]); rho2[i] = lbm2.total_rho(lbm2.lattice[i]);
        if (flag_vect[i] != CellType_twoLayeredPF2D::bulk) {
            ++nwall;
            for (int s = 0; s < 2; ++s)
                for (int k = 0; k < 9; ++k) {
                    const size_t o = (size_t)s * 2 * dim.npop + (size_t)*parity * dim.npop + (size_t)k * n + i;
                    if (lattice[o] != lbm2.lattice[o]) walls_equal = false;
                }
            continue;
        }
        auto u = lbm.velocity(lattice[i]); auto v = lbm2.velocity(lbm2.lattice[i]);
        ux1[i] = u[0]; uy1[i] = u[1]; ux2[i] = v[0]; uy2[i] = v[1];
    }
    E.field("macro_phi(all nodes)", phi2, phi1);
    E.field("total_rho(all nodes)", rho2, rho1);
    E.field("velocity_x", ux2, ux1);
    E.field("velocity_y", uy2, uy1);
    E.exact("bounce_back node populations identical", walls_equal && nwall > 0);
    for (int s = 0; s < 2; ++s) {
        const size_t off = (size_t)s * 2 * dim.npop + (size_t)*parity * dim.npop;
        vector<double> f1(lattice + off, lattice + off + dim.npop), f2(lbm2.lattice + off, lbm2.lattice + off + dim.npop);
        E.field(s ? "populations_g(all nodes)" : "populations_f(all nodes)", f2, f1);
    }
    return E.finish("dropin_hcz_layered2d", n, steps);
}
