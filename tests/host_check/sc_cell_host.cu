// sc_cell_host.cu -- TEST INFRASTRUCTURE (never linked into libclbm.so, never a product path).
//
// The per-cell device functions of the Shan-Chen kernels (multiphase-lbm_b200/csrc/sc_cell.cuh: psi, force sums, collision,
// output fields) compiled FOR THE HOST with -DCLBM_HOST_CHECK, driven by the same two loops as the staged kernels of
// sc_kernels.cu (psi of every node, then gather + collide + push of every bulk node).  tests/test_host_check.py compares
// the result with the oracle at 1e-10, so the arithmetic of a kernel variant can be checked in a container without a GPU
// before it is spent GPU time on.  What this does NOT cover: the kernels' own indexing, shared-memory staging and the
// fused/TMA pipelines -- those are covered by the -m gpu parity tests only.
// The hardware-seeded reciprocal / square root (moments.cuh) fall back to 1/x and sqrt(x) on the host.
#define CLBM_HOST_CHECK 1
#include <cstring>
#include <vector>

#include "../../multiphase-lbm_b200/csrc/mrt.cuh"
#include "../../multiphase-lbm_b200/csrc/sc_cell.cuh"

using namespace clbm;

namespace {

Geom host_geom(const clbm_params *p)
{
    Geom g;
    g.nx = p->nx; g.ny = p->ny; g.nz = p->nz;
    g.G = 0; g.wrapx = 1;                       // single slab, no ghost planes: storage index == reference index
    g.nx_global = p->nx; g.x_offset = 0;
    g.plane = (long long)p->ny * p->nz;
    g.ncs = (long long)p->nx * g.plane;
    return g;
}

template <class L, bool GUO>
void psi_field(const Geom &g, const ModelParams &mp, const double *fin, const uint8_t *flag, double *psi)
{
    for (long long i = 0; i < g.ncs; ++i) {     // body of sc_psi_kernel
        double f[L::Q];
        for (int k = 0; k < L::Q; ++k) f[k] = fin[(size_t)k * g.ncs + i];
        double v = 0.0;
        if (flag[i] != CELL_BB) {
            if (GUO) v = scrt_psi(Mom<L>::sum(f));
            else { bool gp; const double ps = sc_psi_g1(mp, Mom<L>::sum(f), gp); v = gp ? ps : -ps; }
        }
        psi[i] = v;
    }
}

template <class L, bool GUO, bool MRT = false>
void step(const Geom &g, const ModelParams &mp, const double *fin, double *fout, const uint8_t *flag, double *psi)
{
    psi_field<L, GUO>(g, mp, fin, flag, psi);
    for (long long t = 0; t < g.ncs; ++t) {     // body of sc_collide_kernel
        const int x = (int)(t / g.plane), r = (int)(t % g.plane), y = r / g.nz, z = r % g.nz;
        const Nbr n = make_nbr(g, x, y, z);
        if (flag[n.i] != CELL_BULK) continue;
        double f[L::Q], out[L::Q];
        for (int k = 0; k < L::Q; ++k) f[k] = fin[(size_t)k * g.ncs + n.i];
        ScForceSums s = {{0., 0., 0.}, {0., 0., 0.}, 0u};
        sc_gather_force<L, GUO>(s, n, flag, psi);
        const double pc = psi[n.i];
        if constexpr (GUO) scrt_collide<L>(mp, f, s, Mom<L>::sum(f), pc, out);
        else if constexpr (MRT) sc_collide_mrt<L>(mp, f, s, Mom<L>::sum(f), fabs(pc), !std::signbit(pc), out);
        else sc_collide<L>(mp, f, s, fabs(pc), !std::signbit(pc), out);
        for (int k = 0; k < L::Q; ++k) {
            if (k == L::REST) { fout[(size_t)k * g.ncs + n.i] = out[k]; continue; }
            if (s.wall & (1u << k)) fout[(size_t)L::opp(k) * g.ncs + n.i] = out[k];
            else fout[(size_t)k * g.ncs + n.template at<L>(k)] = out[k];
        }
    }
}

template <class L, bool GUO>
void fields(const Geom &g, const ModelParams &mp, const double *fin, const uint8_t *flag, double *psi, double *const o[8])
{
    psi_field<L, GUO>(g, mp, fin, flag, psi);
    for (long long t = 0; t < g.ncs; ++t) {     // body of sc_fields_kernel
        const int x = (int)(t / g.plane), r = (int)(t % g.plane), y = r / g.nz, z = r % g.nz;
        const Nbr n = make_nbr(g, x, y, z);
        double f[L::Q];
        for (int k = 0; k < L::Q; ++k) f[k] = fin[(size_t)k * g.ncs + n.i];
        double rho = Mom<L>::sum(f), pr = 0.0, u[3] = {0., 0., 0.}, F[3] = {0., 0., 0.};
        if (flag[n.i] == CELL_BULK) {
            ScForceSums s = {{0., 0., 0.}, {0., 0., 0.}, 0u};
            sc_gather_force<L, GUO>(s, n, flag, psi);
            if (GUO) scrt_outputs<L>(mp, f, s, rho, pr, u, F);
            else sc_outputs<L>(mp, f, s, rho, pr, u, F);
        }
        const double v[8] = {rho, pr, u[0], u[1], u[2], F[0], F[1], F[2]};
        for (int j = 0; j < 8; ++j) if (o[j]) o[j][t] = v[j];
    }
}

}  // namespace

// lattice / flag / parity in the reference layout (include/clbm.h); Shan-Chen models only
extern "C" int host_check_sc_step(const clbm_params *p, double *lattice, const uint8_t *flag, int *parity, int nsteps)
{
    if (p->model != CLBM_MODEL_SC_D2Q9 && p->model != CLBM_MODEL_SC_D3Q19) return -1;
    const bool guo = p->sc_force == CLBM_SC_FORCE_EXPGUO;
    if (guo && p->model != CLBM_MODEL_SC_D2Q9) return -1;
    const Geom g = host_geom(p);
    ModelParams mp;
    derive_model_params(p, mp);
    const size_t npop = (size_t)(p->model == CLBM_MODEL_SC_D2Q9 ? 9 : 19) * g.ncs;
    std::vector<double> psi(g.ncs);
    for (int s = 0; s < nsteps; ++s) {
        const double *fin = lattice + (size_t)(*parity) * npop;
        double *fout = lattice + (size_t)(1 - *parity) * npop;
        if (p->model == CLBM_MODEL_SC_D3Q19) step<D3Q19, false>(g, mp, fin, fout, flag, psi.data());
        else if (guo) step<D2Q9, true>(g, mp, fin, fout, flag, psi.data());
        else if (p->collision == CLBM_COLLISION_MRT) step<D2Q9, false, true>(g, mp, fin, fout, flag, psi.data());
        else step<D2Q9, false>(g, mp, fin, fout, flag, psi.data());
        *parity = 1 - *parity;
    }
    return 0;
}

// out[8] = {rho, pressure, ux, uy, uz, fx, fy, fz}, NULL entries skipped
extern "C" int host_check_sc_fields(const clbm_params *p, const double *lattice, const uint8_t *flag, int parity, double *const out[8])
{
    if (p->model != CLBM_MODEL_SC_D2Q9 && p->model != CLBM_MODEL_SC_D3Q19) return -1;
    const bool guo = p->sc_force == CLBM_SC_FORCE_EXPGUO;
    if (guo && p->model != CLBM_MODEL_SC_D2Q9) return -1;
    const Geom g = host_geom(p);
    ModelParams mp;
    derive_model_params(p, mp);
    const size_t npop = (size_t)(p->model == CLBM_MODEL_SC_D2Q9 ? 9 : 19) * g.ncs;
    std::vector<double> psi(g.ncs);
    const double *fin = lattice + (size_t)parity * npop;
    if (p->model == CLBM_MODEL_SC_D3Q19) fields<D3Q19, false>(g, mp, fin, flag, psi.data(), out);
    else if (guo) fields<D2Q9, true>(g, mp, fin, flag, psi.data(), out);
    else fields<D2Q9, false>(g, mp, fin, flag, psi.data(), out);
    return 0;
}

// w = M^-1 S M v of csrc/mrt.cuh (the factorised form the HCZ D2Q9 MRT kernels inline); rates = {s_c, s_e, s_eps, s_q, s_nu}
extern "C" void host_check_mrt9(const double *v, const double *rates, double *w)
{
    const MrtRates S = {rates[0], rates[1], rates[2], rates[3], rates[4]};
    mrt9_relax(v, S, w);
}
