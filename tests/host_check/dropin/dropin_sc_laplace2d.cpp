// Drop-in test around SC/apps/laplace2D.h (untouched): the patch of INTEGRATION.md, compiled against the reference header.
#include "dropin_common.h"
#include "laplace2D.h"
int main(int argc, char** argv)
{
    Args A(argc, argv);
    int nx = A.i("nx", 64), ny = A.i("ny", nx), steps = A.i("steps", 1000), threads = A.i("threads", 1);
    double omega = A.d("omega", 1.0), rhol = A.d("rhol", 0.265), rhog = A.d("rhog", 0.038), rho_w = A.d("rho_w", 0.12);
    double a = A.d("a", 1.0), b = A.d("b", 4.0), R = A.d("R", 1.0), TT0 = A.d("TT0", 0.875), gravity = A.d("gravity", 0.0);
    Dim_Laplace2D dim{nx, ny};
    // ---- the reference's own set-up, as Laplace2D() does it (laplace2D.h:439-474) ----
    vector<double> lattice_vect(LBM_Laplace2D::sizeOfLattice(dim.nelem));
    double* lattice = &lattice_vect[0];
    vector<CellType_Laplace2D> flag_vect(dim.nelem);
    CellType_Laplace2D* flag = &flag_vect[0];
    vector<int> parity_vect{0};
    int* parity = &parity_vect[0];
    auto [c, opp, t] = d2q9_constants_Laplace2D();
    LBM_Laplace2D lbm{lattice, flag, parity, &c[0], &opp[0], &t[0], omega, rhol, rhog, rho_w, a, b, R, TT0, 0.0, gravity, dim};
    lbm.TT = lbm.TT0 * (0.3773 * a / (b * R));
    for_each(lattice, lattice + dim.nelem, [&lbm](double& f0) { lbm.iniLattice(f0); });
    inigeom_Laplace2D(lbm);

    // second copy of the host state for the device path, with its own functor object over it
    vector<double> lattice2_vect(lattice_vect);
    vector<int> parity2_vect{*parity};
    LBM_Laplace2D lbm2 = lbm;
    lbm2.lattice = &lattice2_vect[0];
    lbm2.parity = &parity2_vect[0];

    // ---- (B) INTEGRATION.md: hand the aggregate to the device -------------------------------------------------
    clbm_params p{};                         // scalar members of LBM_Laplace2D (laplace2D.h:104-114)
    p.abi_version = CLBM_ABI_VERSION;  p.model = CLBM_MODEL_SC_D2Q9;  p.sc_force = CLBM_SC_FORCE_LAPLACE;
    p.nx = p.nx_global = dim.nx;  p.ny = dim.ny;  p.nz = 1;  p.x_offset = 0;  p.device = -1;  p.fused = 1;
    p.omega = lbm.omega;  p.gravity = lbm.gravity;
    p.rho_w = lbm.rho_w;  p.a = lbm.a;  p.b = lbm.b;  p.R = lbm.R;  p.TT = lbm.TT;
    clbm_ctx* ctx = nullptr;
    DROPIN_CLBM(clbm_create(&p, &ctx));
    DROPIN_CLBM(clbm_upload(ctx, lbm2.lattice, reinterpret_cast<const uint8_t*>(flag), *lbm2.parity));
    // was: for_each(execution::par_unseq, lattice, lattice + dim.nelem, lbm); *parity = 1 - *parity;   (laplace2D.h:506-507)
    DROPIN_CLBM(clbm_step(ctx, steps));
    DROPIN_CLBM(clbm_download_lattice(ctx, lbm2.lattice, lbm2.parity));   // host arrays current again
    // the device-side diagnostics the patched driver may use instead of a download
    double mass_dev = 0, energy_dev = 0;
    DROPIN_CLBM(clbm_reduce(ctx, CLBM_REDUCE_MASS, &mass_dev));
    DROPIN_CLBM(clbm_reduce(ctx, CLBM_REDUCE_ENERGY, &energy_dev));
    DROPIN_CLBM(clbm_destroy(ctx));

    // ---- (A) the reference alone ----
    run_steps(lbm, lattice, dim.nelem, parity, steps, threads);

    // ---- the reference's accessors on both arrays ----
    ErrList E(A.d("tol", 1e-10));
    E.exact("parity", *lbm2.parity == *parity);
    vector<double> r1(dim.nelem), r2(dim.nelem), p1(dim.nelem), p2(dim.nelem), ux1(dim.nelem), ux2(dim.nelem), uy1(dim.nelem), uy2(dim.nelem);
    for (size_t i = 0; i < dim.nelem; ++i) {
        r1[i] = lbm.density(lattice[i]);              r2[i] = lbm2.density(lbm2.lattice[i]);
        p1[i] = lbm.pressure_node(lattice[i]);        p2[i] = lbm2.pressure_node(lbm2.lattice[i]);
        auto u = lbm.u_actual(lattice[i]);            auto v = lbm2.u_actual(lbm2.lattice[i]);
        ux1[i] = u[0]; uy1[i] = u[1]; ux2[i] = v[0]; uy2[i] = v[1];
    }
    E.field("density", r2, r1);
    E.field("pressure_node", p2, p1);
    E.field("u_actual_x", ux2, ux1);
    E.field("u_actual_y", uy2, uy1);
    // the populations themselves, buffer selected by the parity both sides agree on
    vector<double> f1(lattice + (size_t)*parity * dim.npop, lattice + (size_t)*parity * dim.npop + dim.npop);
    vector<double> f2(lbm2.lattice + (size_t)*parity * dim.npop, lbm2.lattice + (size_t)*parity * dim.npop + dim.npop);
    E.field("populations", f2, f1);
    // the reference's own serial diagnostics on the downloaded array vs the device reductions
    const double m_ref = totalMass_Laplace2D(lbm), e_ref = computeEnergy_Laplace2D(lbm);
    E.field("totalMass(downloaded)", {totalMass_Laplace2D(lbm2)}, {m_ref});
    E.field("totalMass(clbm_reduce)", {mass_dev}, {m_ref});
    E.field("computeEnergy(downloaded)", {computeEnergy_Laplace2D(lbm2)}, {e_ref});
    {   // the device reduction sums in another order: 1e-9 of the (tiny) energy is round-off
        const double r = std::fabs(energy_dev - e_ref) / std::fabs(e_ref);
        E.e.emplace_back("computeEnergy(clbm_reduce)", r);
        if (!(r < 1e-8)) E.ok = false;
    }
    return E.finish("dropin_sc_laplace2d", dim.nelem, steps);
}
