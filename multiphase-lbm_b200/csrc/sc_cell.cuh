// sc_cell.cuh -- per-cell Shan-Chen (Yuan-CS) physics shared by the staged and the fused kernels.
//
// Restates, for one bulk node held in registers,
//   force      SC/apps/laplace2D.h:198-242  (CLBM_SC_FORCE_LAPLACE)
//              SC/apps/contactAngle2D.h:248-293 (CLBM_SC_FORCE_CONTACT)
//   u_eq       SC/apps/laplace2D.h:245-251
//   collideBgk SC/apps/laplace2D.h:272-283 and the rest population :301-305
// The D3Q19 form is the composition SURVEY.md 0.1/8c describes (D3Q19 set of
// PF/apps/laplace3D.h:31-55 + the force/BGK of contactAngle2D.h).
#pragma once
#include "clbm_internal.h"
#include "moments.cuh"

namespace clbm {

// neighbour offsets of one cell (wrap already resolved), in cells
struct Nbr {
    long long i;
    long long oxm, oxp, oym, oyp, ozm, ozp;
    CLBM_D long long at(int cx, int cy, int cz) const
    {
        return i + (cx < 0 ? oxm : (cx > 0 ? oxp : 0)) + (cy < 0 ? oym : (cy > 0 ? oyp : 0)) +
               (cz < 0 ? ozm : (cz > 0 ? ozp : 0));
    }
    template <class L> CLBM_D long long at(int k) const { return at(L::cx(k), L::cy(k), L::cz(k)); }
};

CLBM_D Nbr make_nbr(const Geom &g, int x, int y, int z)
{
    Nbr n;
    n.i = g.idx(x, y, z);
    n.oxm = (long long)(g.wx(x - 1) - x) * g.plane;
    n.oxp = (long long)(g.wx(x + 1) - x) * g.plane;
    n.oym = (long long)(g.wy(y - 1) - y) * g.nz;
    n.oyp = (long long)(g.wy(y + 1) - y) * g.nz;
    n.ozm = (long long)(g.wz(z - 1) - z);
    n.ozp = (long long)(g.wz(z + 1) - z);
    return n;
}

struct ScForceSums {
    double ff[3], bb[3];
    unsigned wall;  // bit k set: neighbour in direction k is a bounce_back node
};

// accumulate the k-th neighbour into the force sums (k is a compile-time constant after unrolling)
template <class L> CLBM_D void sc_force_add(ScForceSums &s, int k, bool is_wall, double psi_nb)
{
    const double tk = L::t(k);
    if (is_wall) {
        s.wall |= 1u << k;
        if (L::cx(k)) s.bb[0] += tk * L::cx(k);
        if (L::cy(k)) s.bb[1] += tk * L::cy(k);
        if (L::cz(k)) s.bb[2] += tk * L::cz(k);
    } else {
        if (L::cx(k)) s.ff[0] += tk * L::cx(k) * psi_nb;
        if (L::cy(k)) s.ff[1] += tk * L::cy(k) * psi_nb;
        if (L::cz(k)) s.ff[2] += tk * L::cz(k) * psi_nb;
    }
}

// total force on the node; rho_c is the raw density (not clamped)
template <class L> CLBM_D void sc_force(const ModelParams &mp, const ScForceSums &s, double rho_c, double F[3])
{
    const ScEos eos{mp.R, mp.TT, mp.a};
    const double Zc = eos.Z(rho_c);
    const double G1 = eos.G1_of_Z(rho_c, Zc);
    const double psi_c = eos.psi_of_Z(rho_c, Zc, G1);
    if (mp.sc_force == CLBM_SC_FORCE_CONTACT) {
        if (rho_c <= 0.0) { F[0] = F[1] = F[2] = 0.0; return; }
        const double Zw = eos.Z(mp.rho_w);
        const double val_w = 6.0 * mp.rho_w * (mp.R * mp.TT * Zw - mp.a * mp.rho_w - (1.0 / 3.0)) / G1;
        const double psi_w = (val_w > 0.0) ? sqrt(val_w) : 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) F[d] = -G1 * psi_c * s.ff[d] + (-G1 * psi_c * psi_w * s.bb[d]);
    } else {
        const double psi_w = eos.psi(mp.rho_w);
#pragma unroll
        for (int d = 0; d < 3; ++d) F[d] = -G1 * psi_c * s.ff[d] + (-G1 * psi_c * psi_w * s.bb[d]);
        F[1] += mp.gravity * rho_c;
    }
}

// BGK collision of all Q populations with the tau-shifted equilibrium velocity
template <class L> CLBM_D void sc_collide(const ModelParams &mp, const double *f, const ScForceSums &s, double *out)
{
    const double rho_raw = Mom<L>::sum(f);
    const double rho = fmax(rho_raw, 1e-14);
    double jx, jy, jz, F[3];
    Mom<L>::first(f, jx, jy, jz);
    sc_force<L>(mp, s, rho_raw, F);
    const double omega = mp.omega, tau = 1.0 / omega;
    const double ux = jx / rho + tau * F[0] / rho;
    const double uy = jy / rho + tau * F[1] / rho;
    const double uz = (L::D == 3) ? jz / rho + tau * F[2] / rho : 0.0;
    const double usqr = 1.5 * (ux * ux + uy * uy + uz * uz);
#pragma unroll
    for (int k = 0; k < L::H; ++k) {
        const double ck_u = L::cx(k) * ux + L::cy(k) * uy + L::cz(k) * uz;
        const double eq = rho * L::t(k) * (1. + 3. * ck_u + 4.5 * ck_u * ck_u - usqr);
        const double eqop = eq - 6.0 * rho * L::t(k) * ck_u;
        out[k] = (1. - omega) * f[k] + omega * eq;
        out[L::opp(k)] = (1. - omega) * f[L::opp(k)] + omega * eqop;
    }
    out[L::REST] = (1. - omega) * f[L::REST] + omega * (rho * L::t(L::REST) * (1. - usqr));
}

// output fields of one bulk node: pressure_node (laplace2D.h:308-315) and u_actual (:252-257)
template <class L>
CLBM_D void sc_outputs(const ModelParams &mp, const double *f, const ScForceSums &s, double &rho_raw, double &pr,
                       double u[3])
{
    rho_raw = Mom<L>::sum(f);
    const double rho = fmax(rho_raw, 1e-14);
    double jx, jy, jz, F[3];
    Mom<L>::first(f, jx, jy, jz);
    sc_force<L>(mp, s, rho_raw, F);
    u[0] = jx / rho + 0.5 * F[0] / rho;
    u[1] = jy / rho + 0.5 * F[1] / rho;
    u[2] = (L::D == 3) ? jz / rho + 0.5 * F[2] / rho : 0.0;
    const ScEos eos{mp.R, mp.TT, mp.a};
    const double Zc = eos.Z(rho_raw), G1 = eos.G1_of_Z(rho_raw, Zc), ps = eos.psi_of_Z(rho_raw, Zc, G1);
    pr = (1.0 / 3.0) * rho_raw + (1.0 / 6.0) * G1 * ps * ps;
}

}  // namespace clbm
