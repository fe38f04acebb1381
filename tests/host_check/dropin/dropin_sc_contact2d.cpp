// Drop-in test around SC/apps/contactAngle2D.h (untouched) -- Shan-Chen with bounce-back walls and the contact-angle force, the
// spec source of BASELINE configs[3]: the reference builds the state, the C ABI advances it, the reference's accessors read it.
#include "dropin_common.h"
#include "contactAngle2D.h"
int main(int argc, char** argv)
{
    Args A(argc, argv);
    int nx = A.i("nx", 96), ny = A.i("ny", 48), steps = A.i("steps", 1000), threads = A.i("threads", 1);
    double omega = A.d("omega", 1.0), rhol = A.d("rhol", 0.265), rhog = A.d("rhog", 0.038), rho_w = A.d("rho_w", 0.2);
    double a = A.d("a", 1.0), b = A.d("b", 4.0), R = A.d("R", 1.0), TT0 = A.d("TT0", 0.875), gravity = A.d("gravity", 0.0);
    double RR = A.d("RR", 14.0);
    Dim_contactAngle2D dim{nx, ny};
    // ---- the reference's own set-up, as contactAngle2D() does it (:708-763) ----
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

    vector<double> lattice2_vect(lattice_vect);
    vector<int> parity2_vect{*parity};
    LBM_contactAngle2D lbm2 = lbm;
    lbm2.lattice = &lattice2_vect[0];
    lbm2.parity = &parity2_vect[0];

    // ---- (B) through the C ABI ----
    clbm_params p{};                         // scalar members of LBM_contactAngle2D
    p.abi_version = CLBM_ABI_VERSION;  p.model = CLBM_MODEL_SC_D2Q9;  p.sc_force = CLBM_SC_FORCE_CONTACT;
    p.nx = p.nx_global = dim.nx;  p.ny = dim.ny;  p.nz = 1;  p.x_offset = 0;  p.device = -1;  p.fused = 1;
    p.omega = lbm.omega;  p.gravity = 0.0;   // the contact-angle force has no gravity term (:248-293)
    p.rho_w = lbm.rho_w;  p.a = lbm.a;  p.b = lbm.b;  p.R = lbm.R;  p.TT = lbm.TT;
    clbm_ctx* ctx = nullptr;
    DROPIN_CLBM(clbm_create(&p, &ctx));
    DROPIN_CLBM(clbm_upload(ctx, lbm2.lattice, reinterpret_cast<const uint8_t*>(flag_vect.data()), *lbm2.parity));
    // was: for_each(execution::par_unseq, lattice, lattice + dim.nelem, lbm); *parity = 1 - *parity;   (contactAngle2D.h:801-802)
    DROPIN_CLBM(clbm_step(ctx, steps));
    DROPIN_CLBM(clbm_download_lattice(ctx, lbm2.lattice, lbm2.parity));
    DROPIN_CLBM(clbm_destroy(ctx));

    // ---- (A) the reference alone ----
    run_steps(lbm, lattice, dim.nelem, parity, steps, threads);

    ErrList E(A.d("tol", 1e-10));
    E.exact("parity", *lbm2.parity == *parity);
    const size_t n = dim.nelem;
    size_t walls = 0;
    vector<double> r1(n), r2(n), p1(n, 0.), p2(n, 0.), u1(2 * n, 0.), u2(2 * n, 0.);
    for (size_t i = 0; i < n; ++i) {
        r1[i] = lbm.density(lattice[i]);              r2[i] = lbm2.density(lbm2.lattice[i]);
        if (flag_vect[i] != CellType_contactAngle2D::bulk) { ++walls; continue; }
        p1[i] = lbm.pressure_node(lattice[i]);        p2[i] = lbm2.pressure_node(lbm2.lattice[i]);
        auto u = lbm.u_actual(lattice[i]);            auto v = lbm2.u_actual(lbm2.lattice[i]);
        u1[i] = u[0]; u1[n + i] = u[1]; u2[i] = v[0]; u2[n + i] = v[1];
    }
    E.exact("has_walls", walls >= 2 * (size_t)dim.nx);
    E.field("density", r2, r1);
    E.field("pressure_node", p2, p1);
    E.field("u_actual", u2, u1);             // as a vector
    vector<double> f1(lattice + (size_t)*parity * dim.npop, lattice + (size_t)*parity * dim.npop + dim.npop);
    vector<double> f2(lbm2.lattice + (size_t)*parity * dim.npop, lbm2.lattice + (size_t)*parity * dim.npop + dim.npop);
    E.field("populations", f2, f1);
    return E.finish("dropin_sc_contact2d", n, steps);
}
