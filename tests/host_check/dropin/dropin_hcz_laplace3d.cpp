// Drop-in test around PF/apps/laplace3D.h (untouched) -- the HCZ D3Q19 half of north_star's metric: the reference builds the
// state, the C ABI advances it, the reference's accessors read the result.
#include "dropin_common.h"
#include "laplace3D.h"
int main(int argc, char** argv)
{
    Args A(argc, argv);
    int nx = A.i("nx", 10), ny = A.i("ny", nx), nz = A.i("nz", nx), steps = A.i("steps", 40), threads = A.i("threads", 1);
    double omega = A.d("omega", 1.0), phi_l = A.d("phi_l", 0.251), phi_g = A.d("phi_g", 0.024);
    double rho_l = A.d("rho_l", 0.12), rho_g = A.d("rho_g", 0.04), a = A.d("a", 4.0), b = A.d("b", 4.0);
    double kappa = A.d("kappa", 5e-4), gravity = A.d("gravity", 0.0);
    Dim_laplace3D dim{nx, ny, nz};
    // ---- the reference's own set-up, as laplace3D() does it (:902-923) ----
    vector<CellData> lattice_vect(LBM_laplace3D::sizeOfLattice(dim.nelem));
    CellData* lattice = &lattice_vect[0];
    vector<CellType_laplace3D> flag_vect(dim.nelem);
    vector<int> parity_vect{0};
    int* parity = &parity_vect[0];
    auto [c_vect, opp_vect, t_vect] = d3q19_constants_laplace3D();
    LBM_laplace3D lbm{lattice, &flag_vect[0], parity, &c_vect[0], &opp_vect[0], &t_vect[0],
                      omega, phi_l, phi_g, rho_l, rho_g, a, b, kappa, gravity, dim};
    for_each(lattice, lattice + dim.nelem, [&lbm](CellData& f0) { lbm.iniLattice(f0); });
    inigeom_laplace3D(lbm);

    vector<CellData> lattice2_vect(lattice_vect);
    vector<int> parity2_vect{*parity};
    LBM_laplace3D lbm2 = lbm;
    lbm2.lattice = &lattice2_vect[0];
    lbm2.parity = &parity2_vect[0];

    // ---- (B) through the C ABI ----
    clbm_params p{};                         // scalar members of LBM_laplace3D (laplace3D.h:122-139)
    p.abi_version = CLBM_ABI_VERSION;  p.model = CLBM_MODEL_HCZ_D3Q19;  p.sc_force = CLBM_HCZ_FORCE_GRAVITY;
    p.nx = p.nx_global = dim.nx;  p.ny = dim.ny;  p.nz = dim.nz;  p.x_offset = 0;  p.device = -1;  p.fused = 1;
    p.omega = lbm.omega;  p.gravity = lbm.gravity;
    p.phi_l = lbm.phi_l;  p.phi_g = lbm.phi_g;  p.rho_l = lbm.rho_l;  p.rho_g = lbm.rho_g;  p.a = lbm.a;  p.b = lbm.b;  p.kappa = lbm.kappa;
    clbm_ctx* ctx = nullptr;
    DROPIN_CLBM(clbm_create(&p, &ctx));
    DROPIN_CLBM(clbm_upload(ctx, lbm2.lattice, reinterpret_cast<const uint8_t*>(&flag_vect[0]), *lbm2.parity));
    // was: for_each(execution::par_unseq, lattice, lattice + dim.nelem, lbm); *parity = 1 - *parity;   (laplace3D.h:943-946)
    DROPIN_CLBM(clbm_step(ctx, steps));
    DROPIN_CLBM(clbm_download_lattice(ctx, lbm2.lattice, lbm2.parity));
    DROPIN_CLBM(clbm_destroy(ctx));

    // ---- (A) the reference alone ----
    run_steps(lbm, lattice, dim.nelem, parity, steps, threads);

    ErrList E(A.d("tol", 1e-10));
    E.exact("parity", *lbm2.parity == *parity);
    const size_t n = dim.nelem;
    vector<double> phi1(n), phi2(n), P1(n, 0.), P2(n, 0.), rho1(n), rho2(n), u1(3 * n, 0.), u2(3 * n, 0.);
    for (size_t i = 0; i < n; ++i) {
        auto [ph1, pt1] = lbm.macro_phi_P(lattice[i]);
        auto [ph2, pt2] = lbm2.macro_phi_P(lbm2.lattice[i]);
        phi1[i] = ph1; phi2[i] = ph2;
        rho1[i] = lbm.total_rho(lattice[i]); rho2[i] = lbm2.total_rho(lbm2.lattice[i]);
        if (flag_vect[i] != CellType_laplace3D::bulk) continue;
        P1[i] = lbm.total_P(lattice[i]); P2[i] = lbm2.total_P(lbm2.lattice[i]);
        auto u = lbm.velocity(lattice[i]); auto v = lbm2.velocity(lbm2.lattice[i]);
        for (int d = 0; d < 3; ++d) { u1[d * n + i] = u[d]; u2[d * n + i] = v[d]; }
    }
    E.field("macro_phi", phi2, phi1);
    E.field("total_rho", rho2, rho1);
    E.field("total_P", P2, P1);
    E.field("velocity", u2, u1);             // as a vector: a component may vanish by symmetry
    for (int s = 0; s < 2; ++s) {
        const size_t off = (size_t)s * 2 * dim.npop + (size_t)*parity * dim.npop;
        vector<double> f1(lattice + off, lattice + off + dim.npop), f2(lbm2.lattice + off, lbm2.lattice + off + dim.npop);
        E.field(s ? "populations_g" : "populations_f", f2, f1);
    }
    return E.finish("dropin_hcz_laplace3d", n, steps);
}
