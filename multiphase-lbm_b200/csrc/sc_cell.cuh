// sc_cell.cuh -- per-cell Shan-Chen (Yuan-CS) physics shared by the staged and the fused kernels.
//
// Restates, for one bulk node held in registers,
//   force      SC/apps/laplace2D.h:198-242  (CLBM_SC_FORCE_LAPLACE)
//              SC/apps/contactAngle2D.h:248-293 (CLBM_SC_FORCE_CONTACT)
//   u_eq       SC/apps/laplace2D.h:245-251
//   collideBgk SC/apps/laplace2D.h:272-283 and the rest population :301-305
// The D3Q19 form is the composition SURVEY.md 0.1/8c describes (D3Q19 set of
// PF/apps/laplace3D.h:31-55 + the force/BGK of contactAngle2D.h).
//
// Arithmetic budget (the kernel is FP64-issue sensitive): per node ONE division for
// Z(rho), ONE square root for psi and ONE reciprocal of rho.  The reference's  6(P - rho/3)/G1  with
// G1 = +-1/3 becomes +-18 (P - rho/3); u + tau F/rho becomes (j + tau F)(1/rho); the wall
// pseudopotential (a function of rho_w and the sign of G1 only) is precomputed on the host with the
// reference's own expressions.  All of these differ from the reference by O(1 ulp).
#pragma once
#include "clbm_internal.h"
#include "moments.cuh"
#include "mrt.cuh"

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

// psi(rho) and the sign of G1 with one division and one square root.
//   Z = 1 + (4 rho - 2 rho^2)/(1-rho)^3 ; s = R T Z - a rho - 1/3 ; G1 = sign(s)/3
//   psi = sqrt(6 (P - rho/3)/G1) = sqrt(+-18 (P - rho/3)),  P = rho R T Z - a rho^2   (0 where the radicand <= 0)
CLBM_D double sc_psi_g1(const ModelParams &mp, double rho, bool &g1_pos)
{
    const double d = 1.0 - rho;
    const double Zr = 1.0 + (4.0 * rho - 2.0 * rho * rho) * fast_rcp(d * d * d);
    if (mp.sc_force == CLBM_SC_FORCE_CONSTG) {
        // constant-G mapping (SC/apps/twoLayeredFlow2D.h:183-188): psi^2 = 2 (cs2 rho - (P_eos + p_shift)) / (|G| cs2)
        g1_pos = true;
        const double S = (1.0 / 3.0) * rho - (rho * mp.R * mp.TT * Zr - mp.a * rho * rho + mp.p_shift);
        const double v = mp.kpsi * S;
        return (v > 1e-280) ? fast_sqrt(v) : ((v > 0.0) ? sqrt(v) : 0.0);
    }
    const double s = mp.R * mp.TT * Zr - mp.a * rho - (1.0 / 3.0);
    g1_pos = s > 0.0;
    const double P = rho * mp.R * mp.TT * Zr - mp.a * rho * rho;
    const double q = P - (1.0 / 3.0) * rho;
    const double val = g1_pos ? 18.0 * q : -18.0 * q;
    return (val > 1e-280) ? fast_sqrt(val) : ((val > 0.0) ? sqrt(val) : 0.0);
}

struct ScForceSums {
    double ff[3], bb[3];
    unsigned wall;  // bit k set: neighbour in direction k is a bounce_back node
};

// accumulate the k-th neighbour (k is a compile-time constant after unrolling).  Branch-free for the
// fluid-fluid part; the wall sums are formed afterwards from the mask (rare).
template <class L> CLBM_D void sc_force_add(ScForceSums &s, int k, bool is_wall, double psi_nb)
{
    const double tk = L::t(k);
    const double v = is_wall ? 0.0 : psi_nb;
    if (is_wall) s.wall |= 1u << k;
    if (L::cx(k)) s.ff[0] += tk * L::cx(k) * v;
    if (L::cy(k)) s.ff[1] += tk * L::cy(k) * v;
    if (L::cz(k)) s.ff[2] += tk * L::cz(k) * v;
}

template <class L> CLBM_D void sc_wall_sums(ScForceSums &s)
{
    s.bb[0] = s.bb[1] = s.bb[2] = 0.0;
    if (s.wall == 0u) return;
#pragma unroll
    for (int k = 0; k < L::Q; ++k) {
        if (k == L::REST) continue;
        if (s.wall & (1u << k)) {
            if (L::cx(k)) s.bb[0] += L::t(k) * L::cx(k);
            if (L::cy(k)) s.bb[1] += L::t(k) * L::cy(k);
            if (L::cz(k)) s.bb[2] += L::t(k) * L::cz(k);
        }
    }
}

// total force on the node; rho_c is the raw density (not clamped), psi_c / g1_pos its pseudopotential and branch
template <class L>
CLBM_D void sc_force(const ModelParams &mp, ScForceSums &s, double rho_c, double psi_c, bool g1_pos, double F[3])
{
    sc_wall_sums<L>(s);
    const double G1 = (mp.sc_force == CLBM_SC_FORCE_CONSTG) ? mp.G : (g1_pos ? (1.0 / 3.0) : -(1.0 / 3.0));
    const double psi_w = g1_pos ? mp.psiw_pos : mp.psiw_neg;
    const double a = -G1 * psi_c, b = -G1 * psi_c * psi_w;
#pragma unroll
    for (int d = 0; d < 3; ++d) F[d] = a * s.ff[d] + b * s.bb[d];
    if (mp.sc_force == CLBM_SC_FORCE_CONSTG) {
        // twoLayeredFlow2D.h:224, :256-258: no force on an empty node, else the uniform body force is added as is
        if (rho_c <= 0.0) F[0] = F[1] = F[2] = 0.0;
        else { F[0] += mp.gx; F[1] += mp.gy; }
    } else if (mp.sc_force == CLBM_SC_FORCE_CONTACT) {
        if (rho_c <= 0.0) F[0] = F[1] = F[2] = 0.0;
    } else {
        F[1] += mp.gravity * rho_c;
    }
}

// BGK collision of all Q populations with the tau-shifted equilibrium velocity
// rho_raw = Mom<L>::sum(f), already known to the caller
template <class L>
CLBM_D void sc_collide_rho(const ModelParams &mp, const double *f, ScForceSums &s, double rho_raw, double psi_c, bool g1_pos, double *out)
{
    const double rho = fmax(rho_raw, 1e-14);
    const double inv = fast_rcp(rho);
    double jx, jy, jz, F[3];
    Mom<L>::first(f, jx, jy, jz);
    sc_force<L>(mp, s, rho_raw, psi_c, g1_pos, F);
    const double omega = mp.omega, om1 = 1.0 - omega, tau = mp.tau;
    const double ux = (jx + tau * F[0]) * inv;
    const double uy = (jy + tau * F[1]) * inv;
    const double uz = (L::D == 3) ? (jz + tau * F[2]) * inv : 0.0;
    const double base = 1.0 - 1.5 * (ux * ux + uy * uy + uz * uz);
    const double A = omega * rho;
#pragma unroll
    for (int k = 0; k < L::H; ++k) {
        const double cu = cdot<L>(k, ux, uy, uz);
        const double even = A * L::t(k) * (base + 4.5 * cu * cu);
        const double odd = A * L::t(k) * 3.0 * cu;
        out[k] = om1 * f[k] + (even + odd);
        out[L::opp(k)] = om1 * f[L::opp(k)] + (even - odd);
    }
    out[L::REST] = om1 * f[L::REST] + A * L::t(L::REST) * base;
}

template <class L>
CLBM_D void sc_collide(const ModelParams &mp, const double *f, ScForceSums &s, double psi_c, bool g1_pos, double *out)
{
    sc_collide_rho<L>(mp, f, s, Mom<L>::sum(f), psi_c, g1_pos, out);
}

// MRT relaxation of the Yuan-CS Shan-Chen collision (clbm_params.collision = CLBM_COLLISION_MRT, D2Q9): the same tau-shifted
// equilibrium velocity u + tau F / rho with tau = 1/omega as sc_collide_rho, relaxed in the moment basis of mrt.cuh:
//   out = f - M^-1 S M (f - eq),   S = (omega, s_e, s_eps, omega, s_q, omega, s_q, omega, omega);  S = omega I is BGK.
// The reference's Shan-Chen functors are BGK only: parity of this operator is unpinned against the reference.
template <class L>
CLBM_D void sc_collide_mrt(const ModelParams &mp, const double *f, ScForceSums &s, double rho_raw, double psi_c, bool g1_pos, double *out)
{
    const double rho = fmax(rho_raw, 1e-14);
    const double inv = fast_rcp(rho);
    double jx, jy, jz, F[3];
    Mom<L>::first(f, jx, jy, jz);
    sc_force<L>(mp, s, rho_raw, psi_c, g1_pos, F);
    const double ux = (jx + mp.tau * F[0]) * inv;
    const double uy = (jy + mp.tau * F[1]) * inv;
    const double uz = (L::D == 3) ? (jz + mp.tau * F[2]) * inv : 0.0;
    const double base = 1.0 - 1.5 * (ux * ux + uy * uy + uz * uz);
    double v[L::Q], w[L::Q];
#pragma unroll
    for (int k = 0; k < L::Q; ++k) {
        const double cu = cdot<L>(k, ux, uy, uz);
        v[k] = f[k] - rho * L::t(k) * (base + 3.0 * cu + 4.5 * cu * cu);
    }
    const MrtRates S = {mp.omega, mp.s_e, mp.s_eps, mp.s_q, mp.omega};
    if constexpr (L::Q == 9) mrt9_relax(v, S, w);
    else mrt19_relax(v, S, w);   // D3Q19: the basis of mrt.cuh / oracle mrt19_rows (no reference operator: parity unpinned)
#pragma unroll
    for (int k = 0; k < L::Q; ++k) out[k] = f[k] - w[k];
}

// output fields of one bulk node: pressure_node (laplace2D.h:308-315) and u_actual (:252-257)
template <class L>
CLBM_D void sc_outputs(const ModelParams &mp, const double *f, ScForceSums &s, double &rho_raw, double &pr, double u[3], double F[3])
{
    rho_raw = Mom<L>::sum(f);
    const double rho = fmax(rho_raw, 1e-14);
    double jx, jy, jz;
    Mom<L>::first(f, jx, jy, jz);
    bool g1_pos;
    const double ps = sc_psi_g1(mp, rho_raw, g1_pos);
    sc_force<L>(mp, s, rho_raw, ps, g1_pos, F);
    u[0] = jx / rho + 0.5 * F[0] / rho;
    u[1] = jy / rho + 0.5 * F[1] / rho;
    u[2] = (L::D == 3) ? jz / rho + 0.5 * F[2] / rho : 0.0;
    const double G1 = g1_pos ? (1.0 / 3.0) : -(1.0 / 3.0);
    pr = (1.0 / 3.0) * rho_raw + (1.0 / 6.0) * G1 * ps * ps;
    if (mp.sc_force == CLBM_SC_FORCE_CONSTG) {   // pressure_node = thermodynamic EOS pressure (twoLayeredFlow2D.h:191-194)
        const double d = 1.0 - rho_raw;
        pr = rho_raw * mp.R * mp.TT * (1.0 + (4.0 * rho_raw - 2.0 * rho_raw * rho_raw) / (d * d * d)) - mp.a * rho_raw * rho_raw;
    }
}

// ---- Rayleigh-Taylor variant (CLBM_SC_FORCE_EXPGUO): SC/apps/RayleighTaylor2D.h ----------------------------------------
// A compile-time variant (template flag GUO of the kernels), so that the Yuan-CS instantiations above are untouched.
//   psi = 1 - exp(-rho)                                   :194-196
//   force_ff = -g psi_c sum_k t_k c_k psi(nb), a bounce_back neighbour contributing the psi of the OPPOSITE neighbour
//              (the callers gather that value), + gravity rho in y            :236-289   (force_fw is multiplied by 0: :336-337)
//   u_eq = u + F/(2 rho)                                  :343-351
//   collideBgk with Guo's forcing term                    :370-405, rest population :424-433
CLBM_D double scrt_psi(double rho) { return 1.0 - exp(-rho); }

template <class L>
CLBM_D void scrt_force(const ModelParams &mp, const ScForceSums &s, double rho_c, double psi_c, double F[3])
{
    const double a = -mp.G * psi_c;
#pragma unroll
    for (int d = 0; d < 3; ++d) F[d] = a * s.ff[d];
    F[1] += mp.gravity * rho_c;
}

// omega eq_k + (1 - omega/2) t_k [3 (c_k - u) + 9 (c_k.u) c_k].F  split into the part even in c_k,
//   t_k [A (1 - 1.5 u^2 + 4.5 (c.u)^2) + B (9 (c.u)(c.F) - 3 u.F)],   and the odd part  t_k [3 A (c.u) + 3 B (c.F)],
// A = omega rho, B = 1 - omega/2: the pair (k, opp k) shares both.
template <class L>
CLBM_D void scrt_collide(const ModelParams &mp, const double *f, const ScForceSums &s, double rho_raw, double psi_c, double *out)
{
    const double rho = fmax(rho_raw, 1e-14);
    const double inv = fast_rcp(rho);
    double jx, jy, jz, F[3];
    Mom<L>::first(f, jx, jy, jz);
    scrt_force<L>(mp, s, rho_raw, psi_c, F);
    const double omega = mp.omega, om1 = 1.0 - omega;
    const double ux = (jx + 0.5 * F[0]) * inv;
    const double uy = (jy + 0.5 * F[1]) * inv;
    const double uz = (L::D == 3) ? (jz + 0.5 * F[2]) * inv : 0.0;
    const double base = 1.0 - 1.5 * (ux * ux + uy * uy + uz * uz);
    const double A = omega * rho, B = 1.0 - 0.5 * omega;
    const double uF3 = 3.0 * B * ((L::D == 3) ? (ux * F[0] + uy * F[1] + uz * F[2]) : (ux * F[0] + uy * F[1]));
    const double A3 = 3.0 * A, B3 = 3.0 * B, B9 = 9.0 * B;
#pragma unroll
    for (int k = 0; k < L::H; ++k) {
        const double cu = cdot<L>(k, ux, uy, uz);
        const double cF = cdot<L>(k, F[0], F[1], F[2]);
        const double even = L::t(k) * (A * (base + 4.5 * cu * cu) + (B9 * cu * cF - uF3));
        const double odd = L::t(k) * (A3 * cu + B3 * cF);
        out[k] = om1 * f[k] + (even + odd);
        out[L::opp(k)] = om1 * f[L::opp(k)] + (even - odd);
    }
    out[L::REST] = om1 * f[L::REST] + L::t(L::REST) * (A * base - uF3);
}

// CLBM_COLLISION_MRT for the Rayleigh-Taylor / Guo variant (D2Q9): out = f + F - M^-1 S M (f - eq + F/2) with
//   F_k = t_k [3 (c_k - u) + 9 (c_k.u) c_k] . F,  F_rest = -3 t_rest u.F   (the terms above without their (1 - omega/2) factor;
// oracle scrt_step).  S = omega I is scrt_collide.  The reference functor is BGK: parity unpinned.
template <class L>
CLBM_D void scrt_collide_mrt(const ModelParams &mp, const double *f, const ScForceSums &s, double rho_raw, double psi_c, double *out)
{
    static_assert(L::Q == 9, "the Rayleigh-Taylor variant is D2Q9");
    const double rho = fmax(rho_raw, 1e-14);
    const double inv = fast_rcp(rho);
    double jx, jy, jz, F[3];
    Mom<L>::first(f, jx, jy, jz);
    scrt_force<L>(mp, s, rho_raw, psi_c, F);
    const double ux = (jx + 0.5 * F[0]) * inv;
    const double uy = (jy + 0.5 * F[1]) * inv;
    const double base = 1.0 - 1.5 * (ux * ux + uy * uy);
    const double uF = ux * F[0] + uy * F[1];
    double Fk[9], v[9], w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const double cu = cdot<L>(k, ux, uy, 0.0);
        const double cF = cdot<L>(k, F[0], F[1], 0.0);
        Fk[k] = (k == L::REST) ? L::t(k) * (-3.0 * uF) : L::t(k) * (3.0 * (cF - uF) + 9.0 * cu * cF);
        v[k] = f[k] - rho * L::t(k) * (base + 3.0 * cu + 4.5 * cu * cu) + 0.5 * Fk[k];
    }
    const MrtRates S = {mp.omega, mp.s_e, mp.s_eps, mp.s_q, mp.omega};
    mrt9_relax(v, S, w);
#pragma unroll
    for (int k = 0; k < 9; ++k) out[k] = f[k] + Fk[k] - w[k];
}

// output fields of one bulk node: density, P_eos (:200-208, the Carnahan-Starling pressure with rt = b rho / 4),
// u_eq (what computeEnergy_RayleighTaylor2D :503-516 sums) and force_ff
template <class L>
CLBM_D void scrt_outputs(const ModelParams &mp, const double *f, const ScForceSums &s, double &rho_raw, double &pr, double u[3], double F[3])
{
    rho_raw = Mom<L>::sum(f);
    double jx, jy, jz;
    Mom<L>::first(f, jx, jy, jz);
    scrt_force<L>(mp, s, rho_raw, scrt_psi(rho_raw), F);
    u[0] = jx / rho_raw + F[0] / (2.0 * rho_raw);
    u[1] = jy / rho_raw + F[1] / (2.0 * rho_raw);
    u[2] = (L::D == 3) ? jz / rho_raw + F[2] / (2.0 * rho_raw) : 0.0;
    const double rt = mp.b * rho_raw / 4.0, d = 1.0 - rt;
    pr = (rho_raw / 3.0) * (1.0 + rt + rt * rt - rt * rt * rt) / (d * d * d) - mp.a * rho_raw * rho_raw;
}

// neighbour sums of the force; GUO: a bounce_back neighbour contributes the psi of the opposite neighbour
// (SC/apps/RayleighTaylor2D.h:246-262) instead of the wall-adhesion term
template <class L, bool GUO>
CLBM_D void sc_gather_force(ScForceSums &s, const Nbr &n, const uint8_t *__restrict__ flag, const double *__restrict__ psi)
{
#pragma unroll
    for (int k = 0; k < L::Q; ++k) {
        if (k == L::REST) continue;  // c = 0: contributes nothing, and a bulk node is never its own wall
        const long long nb = n.at<L>(k);
        const bool w = flag[nb] == CELL_BB;
        if constexpr (GUO) {
            if (w) s.wall |= 1u << k;
            sc_force_add<L>(s, k, false, fabs(psi[w ? n.at<L>(L::opp(k)) : nb]));
        } else {
            sc_force_add<L>(s, k, w, w ? 0.0 : fabs(psi[nb]));
        }
    }
}

// Driving force of the two HCZ D2Q9 variants from kappa rho grad(lap X): PF/apps/rayleighTaylor2D.h:325-327 (gravity in y)
// and PF/apps/twoLayeredFlow2D.h:316-317 (rho gx + Gx_const in x, nothing in y)
CLBM_D void hcz2d_force(const ModelParams &mp, double rho, double glx, double gly, double &forcex, double &forcey)
{
    forcex = mp.kappa * rho * glx;
    forcey = mp.kappa * rho * gly;
    if (mp.sc_force == CLBM_HCZ_FORCE_LAYERED) forcex += rho * mp.gx + mp.gx_const;
    else forcey += mp.gravity * rho;
}

}  // namespace clbm
