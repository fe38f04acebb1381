/*
 * clbm_oracle.c -- CPU ORACLE (test infrastructure, NOT a product path).
 *
 * A memoised plain-C restatement of the per-cell functors of the reference
 * (AmooMaD/Multiphase-LBM).  The reference recomputes every macroscopic field
 * recursively per neighbour; here every field is computed
 * once per step into a scratch array with the *same expressions, the same
 * association and the same k = 0..Q-1 summation order*, so results are meant to
 * be bit-identical to the untouched reference functor (checked by
 * tests/test_oracle_vs_reference.py against tests/golden/, which oracle/_ref
 * generated from the reference headers themselves).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file.  The product library (libclbm.so)
 * never links or calls it.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off: the reference is built without FMA contraction
 * (SURVEY.md A.5); we keep IEEE double operations exactly as written.
 *
 * Parity status:
 *   SC D2Q9 (laplace2D.h, contactAngle2D.h, twoLayeredFlow2D.h, RayleighTaylor2D.h), HCZ D2Q9 (rayleighTaylor2D.h, twoLayeredFlow2D.h),
 *   HCZ D3Q19 (laplace3D.h): PINNED against the reference functor (golden fixtures).
 *   SC D3Q19: the reference has no such C++ functor -> no functor fixtures.  Pinned instead (tests/test_sc3d_oracle_symmetry.py,
 *   CPU suite) to the reference's only 3-D Shan-Chen code, the Fortran D3Q19 listing SC/apps/fortran (no Fortran compiler
 *   here: restated in numpy with its own direction ordering): the force of calcu_Fxy (:945-1150) with walls at 1e-13, and
 *   the whole stream / getuv / calcu_Fxy / collision loop on a periodic lattice, populations at 1e-12 after 1000 steps
 *   (the listing's on-node bounce-back differs from the case files' half-way rule, so walls are pinned through the
 *   z-uniform and x-uniform D3Q19 == D2Q9 projections onto the functor-pinned contactAngle2D model instead).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/clbm.h"

#define BB 0 /* CellType::bounce_back */
#define BULK 1

/* ---- lattice constants ---------------------------------------------------
 * D2Q9: SC/apps/laplace2D.h:29-41   D3Q19: PF/apps/laplace3D.h:31-55 */
static const int C9[9][2] = {{-1, 0}, {0, -1}, {-1, -1}, {-1, 1}, {0, 0}, {1, 0}, {0, 1}, {1, 1}, {1, -1}};
static const int OPP9[9] = {5, 6, 7, 8, 4, 0, 1, 2, 3};
static const double T9[9] = {1. / 9., 1. / 9., 1. / 36., 1. / 36., 4. / 9., 1. / 9., 1. / 9., 1. / 36., 1. / 36.};

static const int C19[19][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {-1, -1, 0}, {-1, 1, 0}, {-1, 0, -1}, {-1, 0, 1},
                               {0, -1, -1}, {0, -1, 1}, {0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0},
                               {1, -1, 0}, {1, 0, 1}, {1, 0, -1}, {0, 1, 1}, {0, 1, -1}};
static const int OPP19[19] = {10, 11, 12, 13, 14, 15, 16, 17, 18, 9, 0, 1, 2, 3, 4, 5, 6, 7, 8};
static const double T19[19] = {1. / 18., 1. / 18., 1. / 18., 1. / 36., 1. / 36., 1. / 36., 1. / 36., 1. / 36., 1. / 36.,
                               1. / 3.,
                               1. / 18., 1. / 18., 1. / 18., 1. / 36., 1. / 36., 1. / 36., 1. / 36., 1. / 36., 1. / 36.};

static double dmax(double a, double b) { return a > b ? a : b; }

/* scratch arrays, grown on demand (single caller at a time) */
#define NSCR 16
static double *scr[NSCR];
static size_t scr_n[NSCR];
static double *scratch(int slot, size_t n)
{
    if (scr_n[slot] < n) {
        free(scr[slot]);
        scr[slot] = (double *)malloc(n * sizeof(double));
        scr_n[slot] = n;
    }
    return scr[slot];
}

/* ===========================================================================
 * Shan-Chen, Yuan / Carnahan-Starling EOS
 * SC/apps/laplace2D.h:173-195 == SC/apps/contactAngle2D.h:164-187
 * ======================================================================== */
typedef struct { double R, TT, a; } sc_eos;

static double sc_cs2(void) { return 1.0 / 3.0; }
static double sc_Z(double rho)
{
    const double d = (1.0 - rho);
    const double frac = (4.0 * rho - 2.0 * rho * rho) / (d * d * d);
    return 1.0 + frac;
}
static double sc_P(const sc_eos *e, double rho) { return rho * e->R * e->TT * sc_Z(rho) - e->a * rho * rho; }
static double sc_G1(const sc_eos *e, double rho)
{
    const double s = e->R * e->TT * sc_Z(rho) - e->a * rho - sc_cs2();
    return (s > 0.0) ? sc_cs2() : -sc_cs2();
}
static double sc_psi(const sc_eos *e, double rho)
{
    const double P = sc_P(e, rho);
    const double G1 = sc_G1(e, rho);
    const double val = 6.0 * (P - sc_cs2() * rho) / G1;
    return (val > 0.0) ? sqrt(val) : 0.0;
}

/* constant-G mapping of SC/apps/twoLayeredFlow2D.h:183-188: psi^2 = 2 (cs2 rho - (P_eos + p_shift)) / (|G| cs2) */
static double sc_psi_constg(const sc_eos *e, double rho, double G, double p_shift)
{
    const double P = sc_P(e, rho) + p_shift;
    const double S = sc_cs2() * rho - P;
    if (S <= 0.0) return 0.0;
    return sqrt(2.0 * S / (fabs(G) * sc_cs2()));
}
/* psi of the variant selected by p->sc_force */
static double sc_psi_of(const clbm_params *p, const sc_eos *e, double rho)
{
    return p->sc_force == CLBM_SC_FORCE_CONSTG ? sc_psi_constg(e, rho, p->G, p->p_shift) : sc_psi(e, rho);
}

/* density / raw momentum of one node from the "in" buffer */
static double sc2_density(const double *fin, size_t ne, size_t i)
{ /* SC/apps/laplace2D.h:148-154 */
    double X_M1 = fin[0 * ne + i] + fin[2 * ne + i] + fin[3 * ne + i];
    double X_P1 = fin[5 * ne + i] + fin[7 * ne + i] + fin[8 * ne + i];
    double X_0 = fin[6 * ne + i] + fin[1 * ne + i] + fin[4 * ne + i];
    return X_M1 + X_P1 + X_0;
}
static void sc2_ucommon(const double *fin, size_t ne, size_t i, double u[2])
{ /* SC/apps/laplace2D.h:156-170 */
    double rho = dmax(sc2_density(fin, ne, i), 1e-14);
    double X_M1 = fin[0 * ne + i] + fin[2 * ne + i] + fin[3 * ne + i];
    double X_P1 = fin[5 * ne + i] + fin[7 * ne + i] + fin[8 * ne + i];
    double Y_M1 = fin[1 * ne + i] + fin[2 * ne + i] + fin[8 * ne + i];
    double Y_P1 = fin[3 * ne + i] + fin[6 * ne + i] + fin[7 * ne + i];
    u[0] = (X_P1 - X_M1) / rho;
    u[1] = (Y_P1 - Y_M1) / rho;
}
/* D3Q19 composition: summation groups of PF/apps/laplace3D.h:222-233,244-256 */
static double sc3_density(const double *fin, size_t ne, size_t i)
{
#define F(k) fin[(size_t)(k) * ne + i]
    double X_M1 = F(0) + F(3) + F(4) + F(5) + F(6);
    double X_P1 = F(10) + F(13) + F(14) + F(15) + F(16);
    double X_0 = F(9) + F(1) + F(2) + F(7) + F(8) + F(11) + F(12) + F(17) + F(18);
    return X_M1 + X_P1 + X_0;
}
static void sc3_ucommon(const double *fin, size_t ne, size_t i, double u[3])
{
    double rho = dmax(sc3_density(fin, ne, i), 1e-14);
    double X_M1 = F(0) + F(3) + F(4) + F(5) + F(6);
    double X_P1 = F(10) + F(13) + F(14) + F(15) + F(16);
    double Y_M1 = F(1) + F(3) + F(7) + F(8) + F(14);
    double Y_P1 = F(4) + F(11) + F(13) + F(17) + F(18);
    double Z_M1 = F(2) + F(5) + F(7) + F(16) + F(18);
    double Z_P1 = F(6) + F(8) + F(12) + F(15) + F(17);
#undef F
    u[0] = (X_P1 - X_M1) / rho;
    u[1] = (Y_P1 - Y_M1) / rho;
    u[2] = (Z_P1 - Z_M1) / rho;
}

/*
 * Shan-Chen force at node (x,y[,z]) from the memoised psi field.
 * mode CLBM_SC_FORCE_LAPLACE: SC/apps/laplace2D.h:198-242
 * mode CLBM_SC_FORCE_CONTACT: SC/apps/contactAngle2D.h:248-293
 * D = 2 or 3; 3-D is the composition described in SURVEY.md 0.1 / 8c.
 */
static void sc_force(const clbm_params *p, const sc_eos *e, int D, const double *psi, const uint8_t *flag,
                     double rho_c, int iX, int iY, int iZ, double Fout[3])
{
    const int Q = (D == 2) ? 9 : 19;
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    double sum_ff[3] = {0., 0., 0.}, sum_bb[3] = {0., 0., 0.};
    double G1, psi_c, psi_w;
    Fout[0] = Fout[1] = Fout[2] = 0.0;

    if (p->sc_force == CLBM_SC_FORCE_CONSTG) {      /* SC/apps/twoLayeredFlow2D.h:218-261 */
        if (rho_c <= 0.0) return;
        G1 = p->G;
        psi_c = sc_psi_constg(e, rho_c, p->G, p->p_shift);
        psi_w = sc_psi_constg(e, p->rho_w, p->G, p->p_shift);
    } else if (p->sc_force == CLBM_SC_FORCE_CONTACT) {
        if (rho_c <= 0.0) return;
        G1 = sc_G1(e, rho_c);
        psi_c = sc_psi(e, rho_c);
        const double Zw = sc_Z(p->rho_w);
        const double val_w = 6.0 * p->rho_w * (e->R * e->TT * Zw - e->a * p->rho_w - sc_cs2()) / G1;
        psi_w = (val_w > 0.0) ? sqrt(val_w) : 0.0;
    } else {
        psi_c = sc_psi(e, rho_c);
        G1 = sc_G1(e, rho_c);
        psi_w = sc_psi(e, p->rho_w);
    }

    for (int k = 0; k < Q; ++k) {
        int cx, cy, cz;
        double tk;
        if (D == 2) { cx = C9[k][0]; cy = C9[k][1]; cz = 0; tk = T9[k]; }
        else { cx = C19[k][0]; cy = C19[k][1]; cz = C19[k][2]; tk = T19[k]; }
        int XX = (iX + cx + nx) % nx;
        int YY = (iY + cy + ny) % ny;
        int ZZ = (iZ + cz + nz) % nz;
        size_t nb = (size_t)ZZ + (size_t)nz * ((size_t)YY + (size_t)ny * XX);
        if (flag[nb] == BB) {
            sum_bb[0] += tk * cx;
            sum_bb[1] += tk * cy;
            sum_bb[2] += tk * cz;
        } else {
            double psi_nb = psi[nb];
            sum_ff[0] += tk * cx * psi_nb;
            sum_ff[1] += tk * cy * psi_nb;
            sum_ff[2] += tk * cz * psi_nb;
        }
    }
    if (p->sc_force == CLBM_SC_FORCE_CONSTG) {
        for (int d = 0; d < 3; ++d) Fout[d] = -G1 * psi_c * sum_ff[d] + (-G1 * psi_c * psi_w * sum_bb[d]);
        Fout[0] += p->gx;
        Fout[1] += p->gy;
    } else if (p->sc_force == CLBM_SC_FORCE_CONTACT) {
        for (int d = 0; d < 3; ++d) Fout[d] = -G1 * psi_c * sum_ff[d] + (-G1 * psi_c * psi_w * sum_bb[d]);
    } else {
        for (int d = 0; d < 3; ++d) {
            double Fd = -G1 * psi_c * sum_ff[d];
            Fd += -G1 * psi_c * psi_w * sum_bb[d];
            Fout[d] = Fd;
        }
        Fout[1] += p->gravity * rho_c;
    }
}

static void sc_psi_field(const clbm_params *p, const sc_eos *e, int D, const double *fin, const uint8_t *flag,
                         double *psi, double *rho)
{
    const size_t ne = (size_t)p->nx * p->ny * p->nz;
#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)ne; ++ii) {
        size_t i = (size_t)ii;
        double r = (D == 2) ? sc2_density(fin, ne, i) : sc3_density(fin, ne, i);
        rho[i] = r;
        psi[i] = (flag[i] == BB) ? 0.0 : sc_psi_of(p, e, r);
    }
}

static void mrt9_relax(const double v[9], const double S[9], double w[9]);   /* defined with the HCZ D2Q9 model below */
static void mrt19_relax(const double v[19], const double S[19], double w[19]);
static void mrt19_rates(const clbm_params *p, double S[19]);

/* one Shan-Chen step: operator() of SC/apps/laplace2D.h:285-306 / contactAngle2D.h:333-355 */
static void sc_step(const clbm_params *p, int D, const double *fin, double *fout, const uint8_t *flag)
{
    const int H = (D == 2) ? 4 : 9;
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    const size_t ne = (size_t)nx * ny * nz;
    const sc_eos e = {p->R, p->TT, p->a};
    const double omega = p->omega;
    double *psi = scratch(0, ne), *rhoa = scratch(1, ne);
    sc_psi_field(p, &e, D, fin, flag, psi, rhoa);

#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)ne; ++ii) {
        size_t i = (size_t)ii;
        if (flag[i] != BULK) continue;
        int iX = (int)(i / ((size_t)ny * nz));
        int rem = (int)(i % ((size_t)ny * nz));
        int iY = rem / nz, iZ = rem % nz;

        double rho = dmax(rhoa[i], 1e-14);
        double u[3] = {0., 0., 0.}, F[3], ueq[3];
        if (D == 2) sc2_ucommon(fin, ne, i, u); else sc3_ucommon(fin, ne, i, u);
        sc_force(p, &e, D, psi, flag, rhoa[i], iX, iY, iZ, F);
        const double tau = 1. / omega;
        for (int d = 0; d < 3; ++d) ueq[d] = u[d] + tau * F[d] / rho;
        double usqr = (D == 2) ? 1.5 * (ueq[0] * ueq[0] + ueq[1] * ueq[1])
                               : 1.5 * (ueq[0] * ueq[0] + ueq[1] * ueq[1] + ueq[2] * ueq[2]);

        if (D == 2 && p->collision == CLBM_COLLISION_MRT) {
            /* MRT relaxation of the same equilibrium (tau-shifted velocity, tau = 1/omega): out = f - M^-1 S M (f - eq),
             * S = (omega, s_e, s_eps, omega, s_q, omega, s_q, omega, omega) in the moment basis of
             * CooLBM_MRT_combustion.cpp:313-347; S = omega I is collideBgk.  No reference implementation: parity unpinned. */
            const double S[9] = {omega, p->s_e, p->s_eps, omega, p->s_q, omega, p->s_q, omega, omega};
            double v[9], w[9];
            for (int k = 0; k < 9; ++k) {
                const double ck_u = C9[k][0] * ueq[0] + C9[k][1] * ueq[1];
                v[k] = fin[(size_t)k * ne + i] - rho * T9[k] * (1. + 3. * ck_u + 4.5 * ck_u * ck_u - usqr);
            }
            mrt9_relax(v, S, w);
            for (int k = 0; k < 9; ++k) {
                const double pop_out = fin[(size_t)k * ne + i] - w[k];
                if (k == 4) { fout[(size_t)k * ne + i] = pop_out; continue; }
                int x2 = (iX + C9[k][0] + nx) % nx, y2 = (iY + C9[k][1] + ny) % ny;
                size_t nb = (size_t)y2 + (size_t)ny * x2;
                if (flag[nb] == BB) fout[(size_t)OPP9[k] * ne + i] = pop_out; else fout[(size_t)k * ne + nb] = pop_out;
            }
            continue;
        }
        if (D == 3 && p->collision == CLBM_COLLISION_MRT) {
            /* D3Q19: the same construction in the moment basis of mrt19_rows() below -- out = f - M^-1 S M (f - eq), S = omega I is collideBgk.
             * No reference implementation (the reference has neither a D3Q19 Shan-Chen functor nor a D3Q19 MRT basis): parity unpinned. */
            double S[19], v[19], w[19];
            mrt19_rates(p, S);
            for (int k = 0; k < 19; ++k) {
                const double ck_u = C19[k][0] * ueq[0] + C19[k][1] * ueq[1] + C19[k][2] * ueq[2];
                v[k] = fin[(size_t)k * ne + i] - rho * T19[k] * (1. + 3. * ck_u + 4.5 * ck_u * ck_u - usqr);
            }
            mrt19_relax(v, S, w);
            for (int k = 0; k < 19; ++k) {
                const double pop_out = fin[(size_t)k * ne + i] - w[k];
                if (k == 9) { fout[(size_t)k * ne + i] = pop_out; continue; }
                int x2 = (iX + C19[k][0] + nx) % nx, y2 = (iY + C19[k][1] + ny) % ny, z2 = (iZ + C19[k][2] + nz) % nz;
                size_t nb = (size_t)z2 + (size_t)nz * ((size_t)y2 + (size_t)ny * x2);
                if (flag[nb] == BB) fout[(size_t)OPP19[k] * ne + i] = pop_out; else fout[(size_t)k * ne + nb] = pop_out;
            }
            continue;
        }
        for (int k = 0; k < H; ++k) {
            int cx, cy, cz, ko;
            double tk;
            if (D == 2) { cx = C9[k][0]; cy = C9[k][1]; cz = 0; tk = T9[k]; ko = OPP9[k]; }
            else { cx = C19[k][0]; cy = C19[k][1]; cz = C19[k][2]; tk = T19[k]; ko = OPP19[k]; }
            const double ck_u = (D == 2) ? cx * ueq[0] + cy * ueq[1] : cx * ueq[0] + cy * ueq[1] + cz * ueq[2];
            const double eq = rho * tk * (1. + 3. * ck_u + 4.5 * ck_u * ck_u - usqr);
            const double eqop = eq - 6.0 * rho * tk * ck_u;
            double pop_out = (1. - omega) * fin[(size_t)k * ne + i] + omega * eq;
            double pop_out_opp = (1. - omega) * fin[(size_t)ko * ne + i] + omega * eqop;
            /* stream(i,k) */
            {
                int x2 = (iX + cx + nx) % nx, y2 = (iY + cy + ny) % ny, z2 = (iZ + cz + nz) % nz;
                size_t nb = (size_t)z2 + (size_t)nz * ((size_t)y2 + (size_t)ny * x2);
                if (flag[nb] == BB) fout[(size_t)ko * ne + i] = pop_out; else fout[(size_t)k * ne + nb] = pop_out;
            }
            /* stream(i,opp k) */
            {
                int x2 = (iX - cx + nx) % nx, y2 = (iY - cy + ny) % ny, z2 = (iZ - cz + nz) % nz;
                size_t nb = (size_t)z2 + (size_t)nz * ((size_t)y2 + (size_t)ny * x2);
                if (flag[nb] == BB) fout[(size_t)k * ne + i] = pop_out_opp; else fout[(size_t)ko * ne + nb] = pop_out_opp;
            }
        }
        {
            int k = H;
            double tk = (D == 2) ? T9[k] : T19[k];
            double eq = rho * tk * (1. - usqr);
            fout[(size_t)k * ne + i] = (1. - omega) * fin[(size_t)k * ne + i] + omega * eq;
        }
    }
}

/* SC output fields: density, pressure_node (laplace2D.h:308-315), u_actual (:252-257) */
static void sc_fields(const clbm_params *p, int D, const double *fin, const uint8_t *flag,
                      double *s0, double *s1, double *ux, double *uy, double *uz)
{
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    const size_t ne = (size_t)nx * ny * nz;
    const sc_eos e = {p->R, p->TT, p->a};
    double *psi = scratch(0, ne), *rhoa = scratch(1, ne);
    sc_psi_field(p, &e, D, fin, flag, psi, rhoa);
#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)ne; ++ii) {
        size_t i = (size_t)ii;
        int iX = (int)(i / ((size_t)ny * nz));
        int rem = (int)(i % ((size_t)ny * nz));
        int iY = rem / nz, iZ = rem % nz;
        double r = rhoa[i];
        if (s0) s0[i] = r;
        if (flag[i] != BULK) {
            if (s1) s1[i] = 0.0;
            if (ux) ux[i] = 0.0;
            if (uy) uy[i] = 0.0;
            if (uz) uz[i] = 0.0;
            continue;
        }
        if (s1 && p->sc_force == CLBM_SC_FORCE_CONSTG) {
            s1[i] = sc_P(&e, r);       /* pressure_node = thermodynamic EOS pressure (twoLayeredFlow2D.h:191-194) */
        } else if (s1) {
            double ps = sc_psi(&e, r), G1 = sc_G1(&e, r);
            if (p->sc_force == CLBM_SC_FORCE_CONTACT) s1[i] = sc_cs2() * r + (G1 / 6.0) * ps * ps;
            else s1[i] = (1.0 / 3.0) * r + (1.0 / 6.0) * G1 * ps * ps;
        }
        double rho = dmax(r, 1e-14), u[3] = {0., 0., 0.}, F[3];
        if (D == 2) sc2_ucommon(fin, ne, i, u); else sc3_ucommon(fin, ne, i, u);
        sc_force(p, &e, D, psi, flag, r, iX, iY, iZ, F);
        if (ux) ux[i] = u[0] + 0.5 * F[0] / rho;
        if (uy) uy[i] = u[1] + 0.5 * F[1] / rho;
        if (uz) uz[i] = (D == 3) ? u[2] + 0.5 * F[2] / rho : 0.0;
    }
}

/* ===========================================================================
 * Shan-Chen Rayleigh-Taylor variant (CLBM_SC_FORCE_EXPGUO) -- SC/apps/RayleighTaylor2D.h
 * psi = 1 - exp(-rho) (:194-196), constant coupling g, mirrored psi at wall neighbours
 * (:246-262), Guo forcing with the half-force velocity (:343-351, :370-436).  D2Q9 only.
 * ======================================================================== */
static double scrt_psi(double dens) { return 1 - exp(-dens); }

static double scrt_density(const double *fin, size_t ne, size_t i) { return sc2_density(fin, ne, i); } /* :172-183 */

static void scrt_ucommon(const double *fin, size_t ne, size_t i, double rho, double u[2])
{ /* :186-206 -- no clamp of rho */
    double X_M1 = fin[0 * ne + i] + fin[2 * ne + i] + fin[3 * ne + i];
    double X_P1 = fin[5 * ne + i] + fin[7 * ne + i] + fin[8 * ne + i];
    double Y_M1 = fin[1 * ne + i] + fin[2 * ne + i] + fin[8 * ne + i];
    double Y_P1 = fin[3 * ne + i] + fin[6 * ne + i] + fin[7 * ne + i];
    u[0] = X_P1 - X_M1;
    u[1] = Y_P1 - Y_M1;
    u[0] /= rho;
    u[1] /= rho;
}
static double scrt_Peos(const clbm_params *p, double rho)
{ /* :200-208 */
    double rt = p->b * rho / 4;
    return (rho / 3.) * (1. + rt + rt * rt - rt * rt * rt) / ((1 - rt) * (1 - rt) * (1 - rt)) - p->a * rho * rho;
}
/* force_ff :236-289 ; psi[] holds scrt_psi(density) of EVERY node (wall nodes have density 0 -> psi 0) */
static void scrt_force_ff(const clbm_params *p, const double *psi, const uint8_t *flag, double rho_c, size_t i, int iX, int iY, double F[2])
{
    const int nx = p->nx, ny = p->ny;
    double fx = 0.0, fy = 0.0;
    const double psi_c = psi[i];
    for (int k = 0; k < 9; ++k) {
        int XX = (iX + C9[k][0] + nx) % nx;
        int YY = iY + C9[k][1];
        size_t nb = (size_t)YY + (size_t)ny * XX;
        if (flag[nb] == BB) {
            int XXX = (iX - C9[k][0] + nx) % nx;
            int YYY = iY - C9[k][1];
            size_t nbb = (size_t)YYY + (size_t)ny * XXX;
            double psi_nb = psi[nbb];
            fx += T9[k] * C9[k][0] * psi_nb;
            fy += T9[k] * C9[k][1] * psi_nb;
        } else {
            double psi_nb = psi[nb];
            fx += T9[k] * C9[k][0] * psi_nb;
            fy += T9[k] * C9[k][1] * psi_nb;
        }
    }
    fx *= -p->G * psi_c;
    fy *= -p->G * psi_c;
    fy += p->gravity * rho_c;
    F[0] = fx;
    F[1] = fy;
}
/* force_fw :294-340 -- multiplied by 0. in the reference; kept literally (it only decides the sign of a zero) */
static void scrt_force_fw(const clbm_params *p, const double *psi, const uint8_t *flag, size_t i, int iX, int iY, double F[2])
{
    const int nx = p->nx, ny = p->ny;
    double fx = 0.0, fy = 0.0;
    const double psi_c = psi[i];
    for (int k = 0; k < 9; ++k) {
        int XX = (iX + C9[k][0] + nx) % nx;
        int YY = iY + C9[k][1];
        size_t nb = (size_t)YY + (size_t)ny * XX;
        if (flag[nb] == BB) {
            double psi_nb = psi[nb];
            fx += T9[k] * C9[k][0] * psi_nb;
            fy += T9[k] * C9[k][1] * psi_nb;
        }
    }
    fx *= -p->G * psi_c * 0.;
    fy *= -p->G * psi_c * 0.;
    F[0] = fx;
    F[1] = fy;
}
static void scrt_psi_field(const double *fin, size_t ne, double *psi, double *rho)
{
#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)ne; ++ii) {
        rho[ii] = scrt_density(fin, ne, (size_t)ii);
        psi[ii] = scrt_psi(rho[ii]);
    }
}
/* u_eq :343-351 */
static void scrt_ueq(const double u[2], const double FF[2], const double FW[2], double rho, double ueq[2])
{
    ueq[0] = u[0] + (FF[0] + FW[0]) / (2 * rho);
    ueq[1] = u[1] + (FF[1] + FW[1]) / (2 * rho);
}
/* operator() :407-436 with collideBgk :370-405 and stream :354-367 */
static void scrt_step(const clbm_params *p, const double *fin, double *fout, const uint8_t *flag)
{
    const int nx = p->nx, ny = p->ny;
    const size_t ne = (size_t)nx * ny;
    const double omega = p->omega;
    double *psi = scratch(0, ne), *rhoa = scratch(1, ne);
    scrt_psi_field(fin, ne, psi, rhoa);
#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)ne; ++ii) {
        size_t i = (size_t)ii;
        if (flag[i] != BULK) continue;
        int iX = (int)(i / (size_t)ny), iY = (int)(i % (size_t)ny);
        double rho = rhoa[i], u[2], FF[2], FW[2], ueq[2];
        scrt_ucommon(fin, ne, i, rho, u);
        scrt_force_ff(p, psi, flag, rho, i, iX, iY, FF);
        scrt_force_fw(p, psi, flag, i, iX, iY, FW);
        scrt_ueq(u, FF, FW, rho, ueq);
        double usqreq = 1.5 * (ueq[0] * ueq[0] + ueq[1] * ueq[1]);
        if (p->collision == CLBM_COLLISION_MRT) {
            /* MRT with Guo's forcing: out = f + F - M^-1 S M (f - eq + F/2), F_k = t_k [3 (c_k - u) + 9 (c_k.u) c_k] . FF  (the term of
             * :370-405 without its (1 - omega/2) factor; rest population as the reference writes it, with FW + FF); S = omega I is the
             * collideBgk below.  Parity unpinned (the reference functor is BGK). */
            const double S[9] = {omega, p->s_e, p->s_eps, omega, p->s_q, omega, p->s_q, omega, omega};
            double Fk[9], v[9], w[9];
            for (int k = 0; k < 9; ++k) {
                double ck_ueq = C9[k][0] * ueq[0] + C9[k][1] * ueq[1];
                double eq_f = rho * T9[k] * (1. + 3. * ck_ueq + 4.5 * ck_ueq * ck_ueq - usqreq);
                double e_u_x = C9[k][0] - ueq[0], e_u_y = C9[k][1] - ueq[1];
                if (k == 4) Fk[k] = T9[k] * ((-3. * ueq[0] * (FW[0] + FF[0])) + (-3. * ueq[1] * (FW[1] + FF[1])));
                else Fk[k] = T9[k] * ((3 * e_u_x + 9 * ck_ueq * C9[k][0]) * FF[0] + (3 * e_u_y + 9 * ck_ueq * C9[k][1]) * FF[1]);
                v[k] = fin[(size_t)k * ne + i] - eq_f + 0.5 * Fk[k];
            }
            mrt9_relax(v, S, w);
            for (int k = 0; k < 9; ++k) {
                const double pop_out = fin[(size_t)k * ne + i] + Fk[k] - w[k];
                if (k == 4) { fout[(size_t)k * ne + i] = pop_out; continue; }
                int x2 = (iX + C9[k][0] + nx) % nx, y2 = iY + C9[k][1];
                size_t nb = (size_t)y2 + (size_t)ny * x2;
                if (flag[nb] == BB) fout[(size_t)OPP9[k] * ne + i] = pop_out; else fout[(size_t)k * ne + nb] = pop_out;
            }
            continue;
        }
        for (int k = 0; k < 4; ++k) {
            const int ko = OPP9[k];
            double F_x = FF[0], F_y = FF[1];
            double e_u_x = C9[k][0] - ueq[0];
            double e_u_x_op = C9[ko][0] - ueq[0];
            double e_u_y = C9[k][1] - ueq[1];
            double e_u_y_op = C9[ko][1] - ueq[1];
            double ck_ueq = C9[k][0] * ueq[0] + C9[k][1] * ueq[1];
            double eq_f = rho * T9[k] * (1. + 3. * ck_ueq + 4.5 * ck_ueq * ck_ueq - usqreq);
            double eq_fopp = eq_f - 6. * rho * T9[k] * ck_ueq;
            double total_F = T9[k] * (1 - 0.5 * omega) * ((3 * e_u_x + 9 * ck_ueq * C9[k][0]) * F_x + (3 * e_u_y + 9 * ck_ueq * C9[k][1]) * F_y);
            double total_Fopp = T9[k] * (1 - 0.5 * omega) * ((3 * e_u_x_op - 9 * ck_ueq * C9[ko][0]) * F_x + (3 * e_u_y_op - 9 * ck_ueq * C9[ko][1]) * F_y);
            double pop_out = (1. - omega) * fin[(size_t)k * ne + i] + omega * eq_f + total_F;
            double pop_out_opp = (1. - omega) * fin[(size_t)ko * ne + i] + omega * eq_fopp + total_Fopp;
            {
                int x2 = (iX + C9[k][0] + nx) % nx, y2 = iY + C9[k][1];
                size_t nb = (size_t)y2 + (size_t)ny * x2;
                if (flag[nb] == BB) fout[(size_t)ko * ne + i] = pop_out; else fout[(size_t)k * ne + nb] = pop_out;
            }
            {
                int x2 = (iX + C9[ko][0] + nx) % nx, y2 = iY + C9[ko][1];
                size_t nb = (size_t)y2 + (size_t)ny * x2;
                if (flag[nb] == BB) fout[(size_t)k * ne + i] = pop_out_opp; else fout[(size_t)ko * ne + nb] = pop_out_opp;
            }
        }
        {
            const int k = 4;
            double eq_f = rho * T9[k] * (1. - usqreq);
            double total_F_center = T9[k] * (1 - 0.5 * omega) * ((-3. * ueq[0] * (FW[0] + FF[0])) + (-3. * ueq[1] * (FW[1] + FF[1])));
            fout[(size_t)k * ne + i] = (1. - omega) * fin[(size_t)k * ne + i] + omega * eq_f + total_F_center;
        }
    }
}
/* output fields: density, P_eos, u_eq (the velocity computeEnergy_RayleighTaylor2D :503-516 sums), force_ff */
static void scrt_fields(const clbm_params *p, const double *fin, const uint8_t *flag, double *s0, double *s1, double *ux, double *uy,
                        double *uz, double *fx, double *fy)
{
    const int nx = p->nx, ny = p->ny;
    const size_t ne = (size_t)nx * ny;
    double *psi = scratch(0, ne), *rhoa = scratch(1, ne);
    scrt_psi_field(fin, ne, psi, rhoa);
#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)ne; ++ii) {
        size_t i = (size_t)ii;
        int iX = (int)(i / (size_t)ny), iY = (int)(i % (size_t)ny);
        double r = rhoa[i], u[2] = {0., 0.}, FF[2] = {0., 0.}, FW[2], ueq[2] = {0., 0.}, pr = 0.0;
        if (flag[i] == BULK) {
            pr = scrt_Peos(p, r);
            scrt_ucommon(fin, ne, i, r, u);
            scrt_force_ff(p, psi, flag, r, i, iX, iY, FF);
            scrt_force_fw(p, psi, flag, i, iX, iY, FW);
            scrt_ueq(u, FF, FW, r, ueq);
        }
        if (s0) s0[i] = r;
        if (s1) s1[i] = pr;
        if (ux) ux[i] = ueq[0];
        if (uy) uy[i] = ueq[1];
        if (uz) uz[i] = 0.0;
        if (fx) fx[i] = FF[0];
        if (fy) fy[i] = FF[1];
    }
}

/* ===========================================================================
 * HCZ D2Q9 -- PF/apps/rayleighTaylor2D.h
 * ======================================================================== */
typedef struct {
    double *phi, *Pt, *ur0, *ur1, *rho, *psiphi, *psirho, *lap, *laprho;
} hcz2_f;

static double hcz_pth_minus(double x, double a, double b)
{ /* psi_phi :237-242 / psi_rho :374-379 with x = phi resp. rho */
    double rt = b * x / 4.0;
    double pth = (x / 3.0) * (1 + rt + rt * rt - rt * rt * rt) / pow(1 - rt, 3) - a * x * x;
    return pth - x / 3.0;
}

/* neighbour with the mirror rule of :255-273: wall neighbour -> opposite neighbour */
static size_t hcz2_nb(const clbm_params *p, const uint8_t *flag, int iX, int iY, int k)
{
    const int nx = p->nx, ny = p->ny;
    int ix = iX + C9[k][0], iy = iY + C9[k][1];
    ix = (ix + nx) % nx;
    size_t nb = (size_t)iy + (size_t)ny * ix;
    if (flag[nb] == BB) {
        int ixbb = iX - C9[k][0], iybb = iY - C9[k][1];
        ixbb = (ixbb + nx) % nx;
        nb = (size_t)iybb + (size_t)ny * ixbb;
    }
    return nb;
}

static void hcz2_grad(const clbm_params *p, const uint8_t *flag, const double *X, int iX, int iY, double g[2])
{ /* grad_psi_phi :341-371, grad_psi_rho :385-412, grad_rho :419-446, grad_lap_phi :501-529 */
    double gx = 0.0, gy = 0.0;
    for (int k = 0; k < 9; ++k) {
        double v = X[hcz2_nb(p, flag, iX, iY, k)];
        gx += T9[k] * C9[k][0] * v;
        gy += T9[k] * C9[k][1] * v;
    }
    g[0] = 3.0 * gx;
    g[1] = 3.0 * gy;
}

static void hcz2_fieldsets(const clbm_params *p, const double *fin, const double *gin, const uint8_t *flag, hcz2_f *F)
{
    const int nx = p->nx, ny = p->ny;
    const size_t ne = (size_t)nx * ny;
    F->phi = scratch(0, ne); F->Pt = scratch(1, ne); F->ur0 = scratch(2, ne); F->ur1 = scratch(3, ne);
    F->rho = scratch(4, ne); F->psiphi = scratch(5, ne); F->psirho = scratch(6, ne); F->lap = scratch(7, ne);
    F->laprho = scratch(8, ne);
#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)ne; ++ii) {
        size_t i = (size_t)ii;
#define f(k) fin[(size_t)(k) * ne + i]
#define g(k) gin[(size_t)(k) * ne + i]
        /* macro_phi_P :197-214 */
        double Xf_M1 = f(0) + f(2) + f(3), Xf_P1 = f(5) + f(7) + f(8), Xf_0 = f(6) + f(1) + f(4);
        double Xg_M1 = g(0) + g(2) + g(3), Xg_P1 = g(5) + g(7) + g(8), Xg_0 = g(6) + g(1) + g(4);
        double phi = Xf_M1 + Xf_P1 + Xf_0;
        F->phi[i] = phi;
        F->Pt[i] = Xg_M1 + Xg_P1 + Xg_0;
        /* macro_u :216-230 */
        double Yg_P1 = g(3) + g(7) + g(6), Yg_M1 = g(2) + g(1) + g(8);
        F->ur0[i] = Xg_P1 - Xg_M1;
        F->ur1[i] = Yg_P1 - Yg_M1;
#undef f
#undef g
        /* total_rho :232-235 */
        double rho = p->rho_g + ((phi - p->phi_g) / (p->phi_l - p->phi_g)) * (p->rho_l - p->rho_g);
        F->rho[i] = rho;
        F->psiphi[i] = hcz_pth_minus(phi, p->a, p->b);
        F->psirho[i] = hcz_pth_minus(rho, p->a, p->b);
    }
    /* laplacian_phi :467-495 (bulk nodes only: walls are never evaluated thanks to the mirror rule) */
#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)ne; ++ii) {
        size_t i = (size_t)ii;
        if (flag[i] != BULK) { F->lap[i] = 0.0; continue; }
        int iX = (int)(i / ny), iY = (int)(i % ny);
        double phi_c = F->phi[i], sum = 0.0;
        for (int k = 0; k < 9; ++k) sum += T9[k] * (F->phi[hcz2_nb(p, flag, iX, iY, k)] - phi_c);
        F->lap[i] = 6.0 * sum;
    }
    /* laplacian_rho (PF/apps/twoLayeredFlow2D.h:245-272): only the layered variant's rest population reads it */
    if (p->sc_force == CLBM_HCZ_FORCE_LAYERED) {
#pragma omp parallel for schedule(static)
        for (long long ii = 0; ii < (long long)ne; ++ii) {
            size_t i = (size_t)ii;
            if (flag[i] != BULK) { F->laprho[i] = 0.0; continue; }
            int iX = (int)(i / ny), iY = (int)(i % ny);
            double rho_c = F->rho[i], sum = 0.0;
            for (int k = 0; k < 9; ++k) sum += T9[k] * (F->rho[hcz2_nb(p, flag, iX, iY, k)] - rho_c);
            F->laprho[i] = 6.0 * sum;
        }
    }
}

/* the driving force of the two HCZ D2Q9 variants from kappa rho grad(lap X): rayleighTaylor2D.h:325-327 (gravity in y)
 * and twoLayeredFlow2D.h:316-317 (rho gx + Gx_const in x) */
static void hcz2_force(const clbm_params *p, double rho, const double glap[2], double *forcex, double *forcey)
{
    if (p->sc_force == CLBM_HCZ_FORCE_LAYERED) {
        *forcex = p->kappa * rho * glap[0] + rho * p->gx + p->gx_const;
        *forcey = p->kappa * rho * glap[1];
    } else {
        *forcex = p->kappa * rho * glap[0];
        *forcey = p->kappa * rho * glap[1];
        *forcey += p->gravity * rho;
    }
}

/* ---- MRT relaxation in the moment basis of CooLBM_MRT_combustion.cpp:313-323 (rho, e, eps, jx, qx, jy, qy, pxx, pxy), written for
 * the k-ordering of the SC / PF case headers: row j of M evaluated at c_k.  The rows are mutually orthogonal, so
 * M^-1 = M^T diag(1/|row|^2) (the M_inv table of :326-336).  w = M^-1 S M v.  PARITY UNPINNED: the reference's SC / HCZ functors
 * are BGK; this operator is pinned to them only at S = omega I. */
static void mrt9_rows(double M[9][9], double norm2[9])
{
    for (int k = 0; k < 9; ++k) {
        const double cx = C9[k][0], cy = C9[k][1], c2 = cx * cx + cy * cy;
        M[0][k] = 1.0;
        M[1][k] = -4.0 + 3.0 * c2;
        M[2][k] = 4.0 - 10.5 * c2 + 4.5 * c2 * c2;
        M[3][k] = cx;
        M[4][k] = (-5.0 + 3.0 * c2) * cx;
        M[5][k] = cy;
        M[6][k] = (-5.0 + 3.0 * c2) * cy;
        M[7][k] = cx * cx - cy * cy;
        M[8][k] = cx * cy;
    }
    for (int j = 0; j < 9; ++j) {
        norm2[j] = 0.0;
        for (int k = 0; k < 9; ++k) norm2[j] += M[j][k] * M[j][k];
    }
}
static void mrt9_relax(const double v[9], const double S[9], double w[9])
{
    double M[9][9], n2[9], m[9];
    mrt9_rows(M, n2);                                            /* 81 exact small-integer products: cheap next to the rest */
    for (int j = 0; j < 9; ++j) {
        m[j] = 0.0;
        for (int k = 0; k < 9; ++k) m[j] += M[j][k] * v[k];      /* moment space (:2431-2436) */
        m[j] = S[j] * m[j];                                      /* collision in moment space (:2440-2442) */
    }
    for (int k = 0; k < 9; ++k) {
        w[k] = 0.0;
        for (int j = 0; j < 9; ++j) w[k] += M[j][k] / n2[j] * m[j];   /* back to population space (:2445-2449) */
    }
}

/* ---- D3Q19: the orthogonal moment basis of d'Humieres, Ginzburg, Krafczyk, Lallemand, Luo (Phil. Trans. R. Soc. A 360, 2002), rows
 * evaluated at the c_k of the PF laplace3D.h ordering:
 *   rho | e = 19 c^2 - 30 | eps = (21 c^4 - 53 c^2 + 24) / 2 | j_a = c_a, q_a = (5 c^2 - 9) c_a  (a = x, y, z) |
 *   3 p_xx = 3 c_x^2 - c^2, 3 pi_xx = (3 c^2 - 5)(3 c_x^2 - c^2) | p_ww = c_y^2 - c_z^2, pi_ww = (3 c^2 - 5)(c_y^2 - c_z^2) |
 *   p_xy, p_yz, p_xz | m_x = (c_y^2 - c_z^2) c_x, m_y = (c_z^2 - c_x^2) c_y, m_z = (c_x^2 - c_y^2) c_z.
 * Rates: conserved moments and the stress moments (3 p_xx, p_ww, p_ab) relax with omega (they set the viscosity and keep the
 * BGK hydrodynamics), e with s_e, eps and the fourth-order pi moments with s_eps, the third-order q and m moments with s_q --
 * the D2Q9 assignment (omega, s_e, s_eps, omega, s_q, ...) of CooLBM_MRT_combustion.cpp:339 carried over by moment order.
 * The reference has no D3Q19 MRT: PARITY UNPINNED, pinned only at S = omega I (= collideBgk). */
static void mrt19_rows(double M[19][19], double norm2[19])
{
    for (int k = 0; k < 19; ++k) {
        const double cx = C19[k][0], cy = C19[k][1], cz = C19[k][2], c2 = cx * cx + cy * cy + cz * cz;
        M[0][k] = 1.0;
        M[1][k] = 19.0 * c2 - 30.0;
        M[2][k] = (21.0 * c2 * c2 - 53.0 * c2 + 24.0) / 2.0;
        M[3][k] = cx;
        M[4][k] = (5.0 * c2 - 9.0) * cx;
        M[5][k] = cy;
        M[6][k] = (5.0 * c2 - 9.0) * cy;
        M[7][k] = cz;
        M[8][k] = (5.0 * c2 - 9.0) * cz;
        M[9][k] = 3.0 * cx * cx - c2;
        M[10][k] = (3.0 * c2 - 5.0) * (3.0 * cx * cx - c2);
        M[11][k] = cy * cy - cz * cz;
        M[12][k] = (3.0 * c2 - 5.0) * (cy * cy - cz * cz);
        M[13][k] = cx * cy;
        M[14][k] = cy * cz;
        M[15][k] = cx * cz;
        M[16][k] = (cy * cy - cz * cz) * cx;
        M[17][k] = (cz * cz - cx * cx) * cy;
        M[18][k] = (cx * cx - cy * cy) * cz;
    }
    for (int j = 0; j < 19; ++j) {
        norm2[j] = 0.0;
        for (int k = 0; k < 19; ++k) norm2[j] += M[j][k] * M[j][k];
    }
}
static void mrt19_rates(const clbm_params *p, double S[19])
{
    const double o = p->omega, e = p->s_e, eps = p->s_eps, q = p->s_q;
    const double r[19] = {o, e, eps, o, q, o, q, o, q, o, eps, o, eps, o, o, o, q, q, q};
    for (int j = 0; j < 19; ++j) S[j] = r[j];
}
static void mrt19_relax(const double v[19], const double S[19], double w[19])
{
    double M[19][19], n2[19], m[19];
    mrt19_rows(M, n2);
    for (int j = 0; j < 19; ++j) {
        m[j] = 0.0;
        for (int k = 0; k < 19; ++k) m[j] += M[j][k] * v[k];
        m[j] = S[j] * m[j];
    }
    for (int k = 0; k < 19; ++k) {
        w[k] = 0.0;
        for (int j = 0; j < 19; ++j) w[k] += M[j][k] / n2[j] * m[j];
    }
}
/* exported for the tests: the rows and their squared norms (orthogonality, M^-1 = M^T diag(1/norm2)) */
void oracle_mrt19_rows(double *M361, double *norm2_19)
{
    double M[19][19];
    mrt19_rows(M, norm2_19);
    for (int j = 0; j < 19; ++j) for (int k = 0; k < 19; ++k) M361[j * 19 + k] = M[j][k];
}

/* velocity :316-337 and total_P :452-460 of one bulk node */
static void hcz2_uP(const clbm_params *p, const uint8_t *flag, const hcz2_f *F, size_t i, int iX, int iY,
                    double u[2], double *Ptot, double glap_phi[2])
{
    double rho = F->rho[i];
    hcz2_grad(p, flag, F->lap, iX, iY, glap_phi);
    u[0] = F->ur0[i];
    u[1] = F->ur1[i];
    double forcex, forcey;
    hcz2_force(p, rho, glap_phi, &forcex, &forcey);
    u[0] += forcex / 6.0;
    u[1] += forcey / 6.0;
    u[0] /= (rho / 3.0);
    u[1] /= (rho / 3.0);
    if (Ptot) {
        double gpsi[2];
        hcz2_grad(p, flag, F->rho, iX, iY, gpsi);
        *Ptot = F->Pt[i] - 0.5 * (u[0] * -gpsi[0] / 3. + u[1] * -gpsi[1] / 3.);
    }
}

static void hcz2_step(const clbm_params *p, const double *fin, double *fout, const double *gin, double *gout,
                      const uint8_t *flag)
{
    const int nx = p->nx, ny = p->ny;
    const size_t ne = (size_t)nx * ny;
    const double omega = p->omega;
    hcz2_f F;
    hcz2_fieldsets(p, fin, gin, flag, &F);

#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)ne; ++ii) {
        size_t i = (size_t)ii;
        if (flag[i] != BULK) continue;
        int iX = (int)(i / ny), iY = (int)(i % ny);
        double u[2], P, glap_phi[2], gpsi_rho[2], gpsi_phi[2];
        hcz2_uP(p, flag, &F, i, iX, iY, u, &P, glap_phi); /* operator() :613-623 */
        double phi = F.phi[i], rho = F.rho[i];
        hcz2_grad(p, flag, F.psirho, iX, iY, gpsi_rho);
        hcz2_grad(p, flag, F.psiphi, iX, iY, gpsi_phi);
        double usqr = 1.5 * (u[0] * u[0] + u[1] * u[1]);

        if (p->collision == CLBM_COLLISION_MRT) {
            /* the same equilibria and forcing terms as collideBgk below, without the (1 - omega/2) factor:
             *   out = in + F - M^-1 S M (in - eq + F/2),   S = (omega, s_e, s_eps, omega, s_q, omega, s_q, omega, omega)
             * (relaxation :2440-2442, forcing relaxed with (I - S/2) :2466; S = omega I gives collideBgk back) */
            const double S[9] = {omega, p->s_e, p->s_eps, omega, p->s_q, omega, p->s_q, omega, omega};
            double Ff[9], Fg[9], vf[9], vg[9], wf[9], wg[9];
            double Ex = gpsi_rho[0], Ey = gpsi_rho[1];
            for (int k = 0; k < 9; ++k) {
                double ck_u = C9[k][0] * u[0] + C9[k][1] * u[1];
                double eqf = phi * T9[k] * (1 + 3 * ck_u + 4.5 * ck_u * ck_u - usqr);
                double eqg = T9[k] * (P + (rho / 3.0) * (3 * ck_u + 4.5 * ck_u * ck_u - usqr));
                double e_u_x = C9[k][0] - u[0], e_u_y = C9[k][1] - u[1];
                double forcex, forcey;
                if (k == 4 && p->sc_force == CLBM_HCZ_FORCE_LAYERED) {   /* rest population: grad lap RHO (B.9) */
                    double glap_rho[2];
                    hcz2_grad(p, flag, F.laprho, iX, iY, glap_rho);
                    hcz2_force(p, rho, glap_rho, &forcex, &forcey);
                } else {
                    hcz2_force(p, rho, glap_phi, &forcex, &forcey);
                }
                Fg[k] = (e_u_x * forcex + e_u_y * forcey) * eqf / phi + ((e_u_x * -Ex) + (e_u_y * -Ey)) * (eqf / phi - T9[k]);
                if (k == 4)   /* the reference writes the rest term with (u.(-E)), not ((0-u).(-E)) (:654-656): kept */
                    Fg[k] = -(u[0] * forcex + u[1] * forcey) * eqf / phi + ((u[0] * -Ex + u[1] * -Ey) * (eqf / phi - T9[k]));
                Ff[k] = ((e_u_x * -gpsi_phi[0]) + (e_u_y * -gpsi_phi[1])) * 3.0 * eqf / phi;
                vf[k] = fin[(size_t)k * ne + i] - eqf + 0.5 * Ff[k];
                vg[k] = gin[(size_t)k * ne + i] - eqg + 0.5 * Fg[k];
            }
            mrt9_relax(vf, S, wf);
            mrt9_relax(vg, S, wg);
            for (int k = 0; k < 9; ++k) {
                double pf = fin[(size_t)k * ne + i] + Ff[k] - wf[k];
                double pg = gin[(size_t)k * ne + i] + Fg[k] - wg[k];
                if (k == 4) { fout[(size_t)k * ne + i] = pf; gout[(size_t)k * ne + i] = pg; continue; }
                int XX = (iX + C9[k][0] + nx) % nx, YY = iY + C9[k][1];
                size_t nb = (size_t)YY + (size_t)ny * XX;
                if (flag[nb] == BB) { fout[(size_t)OPP9[k] * ne + i] = pf; gout[(size_t)OPP9[k] * ne + i] = pg; }
                else { fout[(size_t)k * ne + nb] = pf; gout[(size_t)k * ne + nb] = pg; }
            }
            continue;
        }

        for (int k = 0; k < 4; ++k) { /* collideBgk :552-606 */
            const int ko = OPP9[k];
            double ck_u = C9[k][0] * u[0] + C9[k][1] * u[1];
            double ck_u_op = C9[ko][0] * u[0] + C9[ko][1] * u[1];
            double eqf = phi * T9[k] * (1 + 3 * ck_u + 4.5 * ck_u * ck_u - usqr);
            double eqf_op = phi * T9[ko] * (1 + 3 * ck_u_op + 4.5 * ck_u_op * ck_u_op - usqr);
            double eqg = T9[k] * (P + (rho / 3.0) * (3 * ck_u + 4.5 * ck_u * ck_u - usqr));
            double eqg_op = T9[ko] * (P + (rho / 3.0) * (3 * ck_u_op + 4.5 * ck_u_op * ck_u_op - usqr));
            double e_u_x = C9[k][0] - u[0], e_u_x_op = C9[ko][0] - u[0];
            double e_u_y = C9[k][1] - u[1], e_u_y_op = C9[ko][1] - u[1];
            double forcex, forcey;
            hcz2_force(p, rho, glap_phi, &forcex, &forcey);
            double Ex = gpsi_rho[0], Ey = gpsi_rho[1];
            double fg = (1. - 0.5 * omega) * ((e_u_x * forcex + e_u_y * forcey) * eqf / phi)
                      + (1. - 0.5 * omega) * ((e_u_x * -Ex) + (e_u_y * -Ey)) * (eqf / phi - T9[k]);
            double fg_op = (1. - 0.5 * omega) * ((e_u_x_op * forcex + e_u_y_op * forcey) * eqf_op / phi)
                         + (1. - 0.5 * omega) * ((e_u_x_op * -Ex) + (e_u_y_op * -Ey)) * (eqf_op / phi - T9[k]);
            double ff = (1. - 0.5 * omega) * ((e_u_x * -gpsi_phi[0]) + (e_u_y * -gpsi_phi[1])) * 3.0 * eqf / phi;
            double ff_op = (1. - 0.5 * omega) * ((e_u_x_op * -gpsi_phi[0]) + (e_u_y_op * -gpsi_phi[1])) * 3.0 * eqf_op / phi;
            double pf = (1. - omega) * fin[(size_t)k * ne + i] + omega * eqf + ff;
            double pg = (1. - omega) * gin[(size_t)k * ne + i] + omega * eqg + fg;
            double pf_op = (1. - omega) * fin[(size_t)ko * ne + i] + omega * eqf_op + ff_op;
            double pg_op = (1. - omega) * gin[(size_t)ko * ne + i] + omega * eqg_op + fg_op;
            /* stream :533-549 : x periodic, y not wrapped */
            for (int s = 0; s < 2; ++s) {
                int kk = s ? ko : k, kko = s ? k : ko;
                int XX = (iX + C9[kk][0] + nx) % nx, YY = iY + C9[kk][1];
                size_t nb = (size_t)YY + (size_t)ny * XX;
                double vf = s ? pf_op : pf, vg = s ? pg_op : pg;
                if (flag[nb] == BB) { fout[(size_t)kko * ne + i] = vf; gout[(size_t)kko * ne + i] = vg; }
                else { fout[(size_t)kk * ne + nb] = vf; gout[(size_t)kk * ne + nb] = vg; }
            }
        }
        { /* rest population :642-663 */
            int k = 4;
            double eqf0 = phi * T9[k] * (1. - usqr);
            double eqg0 = T9[k] * (P - (rho / 3.0) * usqr);
            double forcex, forcey;
            if (p->sc_force == CLBM_HCZ_FORCE_LAYERED) { /* rest population: grad lap RHO (twoLayeredFlow2D.h:595-598, B.9) */
                double glap_rho[2];
                hcz2_grad(p, flag, F.laprho, iX, iY, glap_rho);
                hcz2_force(p, rho, glap_rho, &forcex, &forcey);
            } else {
                hcz2_force(p, rho, glap_phi, &forcex, &forcey);
            }
            double Ex = gpsi_rho[0], Ey = gpsi_rho[1];
            double fg0 = (1. - 0.5 * omega) *
                         (-(u[0] * forcex + u[1] * forcey) * eqf0 / phi + ((u[0] * -Ex + u[1] * -Ey) * (eqf0 / phi - T9[k])));
            double ff0 = (1. - 0.5 * omega) * (-3.0 * (u[0] * -gpsi_phi[0] + u[1] * -gpsi_phi[1]) * eqf0 / phi);
            fout[(size_t)k * ne + i] = (1 - omega) * fin[(size_t)k * ne + i] + omega * eqf0 + ff0;
            gout[(size_t)k * ne + i] = (1 - omega) * gin[(size_t)k * ne + i] + omega * eqg0 + fg0;
        }
    }
}

static void hcz2_fields(const clbm_params *p, const double *fin, const double *gin, const uint8_t *flag,
                        double *s0, double *s1, double *s2, double *ux, double *uy)
{
    const int ny = p->ny;
    const size_t ne = (size_t)p->nx * ny;
    hcz2_f F;
    hcz2_fieldsets(p, fin, gin, flag, &F);
#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)ne; ++ii) {
        size_t i = (size_t)ii;
        if (s0) s0[i] = F.phi[i];
        if (s2) s2[i] = F.rho[i];
        if (flag[i] != BULK) {
            if (s1) s1[i] = 0.0;
            if (ux) ux[i] = 0.0;
            if (uy) uy[i] = 0.0;
            continue;
        }
        double u[2], P, gl[2];
        hcz2_uP(p, flag, &F, i, (int)(i / ny), (int)(i % ny), u, &P, gl);
        if (s1) s1[i] = P;
        if (ux) ux[i] = u[0];
        if (uy) uy[i] = u[1];
    }
}

/* ===========================================================================
 * HCZ D3Q19 -- PF/apps/laplace3D.h
 * ======================================================================== */
typedef struct {
    double *phi, *Pt, *ur0, *ur1, *ur2, *rho, *psiphi, *lap;   /* level 0/1 */
    double *gl0, *gl1, *gl2, *gp0, *gp1, *gp2;                /* grad lap phi, grad psi(phi) */
    double *u0, *u1, *u2, *psirho;                            /* level 2 */
} hcz3_f;

static size_t hcz3_nbidx(const clbm_params *p, int iX, int iY, int iZ, int k)
{
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    int ix = (iX + C19[k][0] + nx) % nx, iy = (iY + C19[k][1] + ny) % ny, iz = (iZ + C19[k][2] + nz) % nz;
    return (size_t)iz + (size_t)nz * ((size_t)iy + (size_t)ny * ix);
}

/* gradients with the "wall neighbour -> centre value" fallback (:435-465, :470-500, :506-536) */
static void hcz3_grad(const clbm_params *p, const uint8_t *flag, const double *X, size_t i, int iX, int iY, int iZ,
                      double g[3])
{
    double gx = 0.0, gy = 0.0, gz = 0.0;
    for (int k = 0; k < 19; ++k) {
        size_t nb = hcz3_nbidx(p, iX, iY, iZ, k);
        double v = (flag[nb] == BB) ? X[i] : X[nb];
        gx += T19[k] * C19[k][0] * v;
        gy += T19[k] * C19[k][1] * v;
        gz += T19[k] * C19[k][2] * v;
    }
    g[0] = 3.0 * gx;
    g[1] = 3.0 * gy;
    g[2] = 3.0 * gz;
}

static void hcz3_fieldsets(const clbm_params *p, const double *fin, const double *gin, const uint8_t *flag, hcz3_f *F)
{
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    const size_t ne = (size_t)nx * ny * nz;
    double **slots[] = {&F->phi, &F->Pt, &F->ur0, &F->ur1, &F->ur2, &F->rho, &F->psiphi, &F->lap,
                        &F->gl0, &F->gl1, &F->gl2, &F->gp0, &F->gp1, &F->gp2, &F->u0, &F->u1};
    for (int s = 0; s < 16; ++s) *slots[s] = scratch(s, ne);
    /* two more arrays reuse level-0 slots that are dead by then: ur2 stays, so allocate separately */
    static double *extra[2]; static size_t extra_n;
    if (extra_n < ne) { free(extra[0]); free(extra[1]); extra[0] = malloc(ne * 8); extra[1] = malloc(ne * 8); extra_n = ne; }
    F->u2 = extra[0]; F->psirho = extra[1];

#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)ne; ++ii) {
        size_t i = (size_t)ii;
#define f(k) fin[(size_t)(k) * ne + i]
#define g(k) gin[(size_t)(k) * ne + i]
        /* macro_phi_P :216-237 */
        double Xg_M1 = g(0) + g(3) + g(4) + g(5) + g(6);
        double Xg_P1 = g(10) + g(13) + g(14) + g(15) + g(16);
        double Xg_0 = g(9) + g(1) + g(2) + g(7) + g(8) + g(11) + g(12) + g(17) + g(18);
        double Xf_M1 = f(0) + f(3) + f(4) + f(5) + f(6);
        double Xf_P1 = f(10) + f(13) + f(14) + f(15) + f(16);
        double Xf_0 = f(9) + f(1) + f(2) + f(7) + f(8) + f(11) + f(12) + f(17) + f(18);
        double phi = Xf_M1 + Xf_P1 + Xf_0;
        F->phi[i] = phi;
        F->Pt[i] = Xg_M1 + Xg_P1 + Xg_0;
        /* macro_u :239-258 */
        double Yg_M1 = g(1) + g(3) + g(7) + g(8) + g(14);
        double Yg_P1 = g(4) + g(11) + g(13) + g(17) + g(18);
        double Zg_M1 = g(2) + g(5) + g(7) + g(16) + g(18);
        double Zg_P1 = g(6) + g(8) + g(12) + g(15) + g(17);
        F->ur0[i] = Xg_P1 - Xg_M1;
        F->ur1[i] = Yg_P1 - Yg_M1;
        F->ur2[i] = Zg_P1 - Zg_M1;
#undef f
#undef g
        F->rho[i] = p->rho_g + ((phi - p->phi_g) / (p->phi_l - p->phi_g)) * (p->rho_l - p->rho_g); /* :261-266 */
        { /* psi_phi :268-275 */
            double rt = p->b * phi / 4.0;
            double pth = (phi / 3.0) * (1 + rt + rt * rt - rt * rt * rt) / pow(1 - rt, 3) - p->a * phi * phi;
            F->psiphi[i] = pth - phi / 3.0;
        }
    }
    /* laplacian_phi :370-393 (wall neighbours skipped) */
#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)ne; ++ii) {
        size_t i = (size_t)ii;
        int iX = (int)(i / ((size_t)ny * nz)), rem = (int)(i % ((size_t)ny * nz)), iY = rem / nz, iZ = rem % nz;
        double phi_c = F->phi[i], sum = 0.0;
        for (int k = 0; k < 19; ++k) {
            size_t nb = hcz3_nbidx(p, iX, iY, iZ, k);
            if (flag[nb] != BB) sum += T19[k] * (F->phi[nb] - phi_c);
        }
        F->lap[i] = 6.0 * sum;
    }
    /* grad_lap_phi, grad_psi_phi, velocity :280-312, total_P :318-328, psi_rho :330-336 */
#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)ne; ++ii) {
        size_t i = (size_t)ii;
        int iX = (int)(i / ((size_t)ny * nz)), rem = (int)(i % ((size_t)ny * nz)), iY = rem / nz, iZ = rem % nz;
        double gl[3], gp[3], u[3];
        hcz3_grad(p, flag, F->lap, i, iX, iY, iZ, gl);
        hcz3_grad(p, flag, F->psiphi, i, iX, iY, iZ, gp);
        F->gl0[i] = gl[0]; F->gl1[i] = gl[1]; F->gl2[i] = gl[2];
        F->gp0[i] = gp[0]; F->gp1[i] = gp[1]; F->gp2[i] = gp[2];
        double rho = F->rho[i], phi = F->phi[i];
        u[0] = F->ur0[i]; u[1] = F->ur1[i]; u[2] = F->ur2[i];
        double forcex = p->kappa * phi * gl[0];
        double forcey = p->kappa * phi * gl[1];
        double forcez = p->kappa * phi * gl[2];
        (void)forcez;
        forcey += p->gravity * rho;
        u[0] += forcex / 6.;
        u[1] += forcey / 6.;
        u[2] += forcey / 6.; /* sic: forcey, SURVEY.md B.5 (laplace3D.h:304) */
        u[0] /= (rho / 3.);
        u[1] /= (rho / 3.);
        u[2] /= (rho / 3.);
        F->u0[i] = u[0]; F->u1[i] = u[1]; F->u2[i] = u[2];
        double P = F->Pt[i] - 0.5 * (u[0] * gp[0] + u[1] * gp[1] + u[2] * gp[2]);
        F->psirho[i] = P - rho / 3.0;
    }
}

static void hcz3_step(const clbm_params *p, const double *fin, double *fout, const double *gin, double *gout,
                      const uint8_t *flag)
{
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    const size_t ne = (size_t)nx * ny * nz;
    const double omega = p->omega, kappa = p->kappa, gravity = p->gravity;
    hcz3_f F;
    hcz3_fieldsets(p, fin, gin, flag, &F);

#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)ne; ++ii) {
        size_t i = (size_t)ii;
        if (flag[i] != BULK) continue;
        int iX = (int)(i / ((size_t)ny * nz)), rem = (int)(i % ((size_t)ny * nz)), iY = rem / nz, iZ = rem % nz;
        double u[3] = {F.u0[i], F.u1[i], F.u2[i]};
        double phi = F.phi[i], rho = F.rho[i];
        double grad_psi_phi[3] = {F.gp0[i], F.gp1[i], F.gp2[i]};
        double grad_lap_phi[3] = {F.gl0[i], F.gl1[i], F.gl2[i]};
        double P = F.Pt[i] - 0.5 * (u[0] * grad_psi_phi[0] + u[1] * grad_psi_phi[1] + u[2] * grad_psi_phi[2]);
        double grad_psi_rho[3];
        hcz3_grad(p, flag, F.psirho, i, iX, iY, iZ, grad_psi_rho);
        double forcex = kappa * phi * grad_lap_phi[0];
        double forcey = kappa * phi * grad_lap_phi[1];
        double forcez = kappa * phi * grad_lap_phi[2];
        forcey += gravity * rho;
        double Ex = grad_psi_rho[0], Ey = grad_psi_rho[1], Ez = grad_psi_rho[2];
        double usqr = 1.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);

        if (p->collision == CLBM_COLLISION_MRT) {
            /* the equilibria and forcing terms of collideBgk below without the (1 - omega/2) factor, relaxed in the D3Q19 moment basis:
             *   out = in + F - M^-1 S M (in - eq + F/2);  S = omega I gives collideBgk back.  Parity unpinned (see mrt19_rows). */
            double S[19], Ff[19], Fg[19], vf[19], vg[19], wf[19], wg[19];
            mrt19_rates(p, S);
            for (int k = 0; k < 19; ++k) {
                double ck_u = C19[k][0] * u[0] + C19[k][1] * u[1] + C19[k][2] * u[2];
                double eqf = phi * T19[k] * (1 + 3 * ck_u + 4.5 * ck_u * ck_u - usqr);
                double eqg = T19[k] * (P + (rho / 3.0) * (3 * ck_u + 4.5 * ck_u * ck_u - usqr));
                double e_u_x = C19[k][0] - u[0], e_u_y = C19[k][1] - u[1], e_u_z = C19[k][2] - u[2];
                Fg[k] = (e_u_x * forcex + e_u_y * forcey + e_u_z * forcez) * eqf / phi
                      + ((e_u_x * -1 * Ex) + (e_u_y * -1 * Ey) + (e_u_z * -1 * Ez)) * (eqf / phi - T19[k]);
                Ff[k] = ((e_u_x * -1 * grad_psi_phi[0]) + (e_u_y * -1 * grad_psi_phi[1]) + (e_u_z * -1 * grad_psi_phi[2])) * 3. * eqf / rho;
                vf[k] = fin[(size_t)k * ne + i] - eqf + 0.5 * Ff[k];
                vg[k] = gin[(size_t)k * ne + i] - eqg + 0.5 * Fg[k];
            }
            mrt19_relax(vf, S, wf);
            mrt19_relax(vg, S, wg);
            for (int k = 0; k < 19; ++k) {
                double pf = fin[(size_t)k * ne + i] + Ff[k] - wf[k];
                double pg = gin[(size_t)k * ne + i] + Fg[k] - wg[k];
                if (k == 9) { fout[(size_t)k * ne + i] = pf; gout[(size_t)k * ne + i] = pg; continue; }
                size_t nb = hcz3_nbidx(p, iX, iY, iZ, k);
                if (flag[nb] == BB) { fout[(size_t)OPP19[k] * ne + i] = pf; gout[(size_t)OPP19[k] * ne + i] = pg; }
                else { fout[(size_t)k * ne + nb] = pf; gout[(size_t)k * ne + nb] = pg; }
            }
            continue;
        }
        for (int k = 0; k < 9; ++k) { /* collideBgk :562-624 */
            const int ko = OPP19[k];
            double ck_u = C19[k][0] * u[0] + C19[k][1] * u[1] + C19[k][2] * u[2];
            double eqf = phi * T19[k] * (1 + 3 * ck_u + 4.5 * ck_u * ck_u - usqr);
            double eqf_op = eqf - 6 * phi * T19[k] * ck_u;
            double eqg = T19[k] * (P + (rho / 3.0) * (3 * ck_u + 4.5 * ck_u * ck_u - usqr));
            double eqg_op = eqg - 6 * (rho / 3.0) * T19[k] * ck_u;
            double e_u_x = C19[k][0] - u[0], e_u_x_op = C19[ko][0] - u[0];
            double e_u_y = C19[k][1] - u[1], e_u_y_op = C19[ko][1] - u[1];
            double e_u_z = C19[k][2] - u[2], e_u_z_op = C19[ko][2] - u[2];
            double fg = (1. - 0.5 * omega) * ((e_u_x * forcex + e_u_y * forcey + e_u_z * forcez) * eqf / phi)
                      + (1. - 0.5 * omega) * ((e_u_x * -1 * Ex) + (e_u_y * -1 * Ey) + (e_u_z * -1 * Ez)) * (eqf / phi - T19[k]);
            double ff = (1. - 0.5 * omega) * ((e_u_x * -1 * grad_psi_phi[0]) + (e_u_y * -1 * grad_psi_phi[1]) + (e_u_z * -1 * grad_psi_phi[2])) * 3. * eqf / rho;
            double fg_op = (1. - 0.5 * omega) * ((e_u_x_op * forcex + e_u_y_op * forcey + e_u_z_op * forcez) * eqf_op / phi)
                         + (1. - 0.5 * omega) * ((e_u_x_op * -1 * Ex + e_u_y_op * -1 * Ey + e_u_z_op * -1 * Ez)) * (eqf_op / phi - T19[k]);
            double ff_op = (1. - 0.5 * omega) * ((e_u_x_op * -1 * grad_psi_phi[0]) + (e_u_y_op * -1 * grad_psi_phi[1]) + (e_u_z_op * -1 * grad_psi_phi[2])) * 3. * eqf_op / rho;
            double pf = (1. - omega) * fin[(size_t)k * ne + i] + omega * eqf + ff;
            double pg = (1. - omega) * gin[(size_t)k * ne + i] + omega * eqg + fg;
            double pf_op = (1. - omega) * fin[(size_t)ko * ne + i] + omega * eqf_op + ff_op;
            double pg_op = (1. - omega) * gin[(size_t)ko * ne + i] + omega * eqg_op + fg_op;
            for (int s = 0; s < 2; ++s) { /* stream :539-559 */
                int kk = s ? ko : k, kko = s ? k : ko;
                size_t nb = hcz3_nbidx(p, iX, iY, iZ, kk);
                double vf = s ? pf_op : pf, vg = s ? pg_op : pg;
                if (flag[nb] == BB) { fout[(size_t)kko * ne + i] = vf; gout[(size_t)kko * ne + i] = vg; }
                else { fout[(size_t)kk * ne + nb] = vf; gout[(size_t)kk * ne + nb] = vg; }
            }
        }
        { /* rest population :664-677 */
            int k = 9;
            double eqf0 = phi * T19[k] * (1. - usqr);
            double eqg0 = T19[k] * (P - (rho / 3.0) * usqr);
            double fg0 = (1. - 0.5 * omega) * -1 * (u[0] * forcex + u[1] * forcey + u[2] * forcez) * eqf0 / phi
                       + (1. - 0.5 * omega) * -1 * (u[0] * -1 * Ex + u[1] * -1 * Ey + u[2] * -1 * Ez) * (eqf0 / phi - T19[k]);
            double ff0 = (1. - 0.5 * omega) * -3. * eqf0 * (u[0] * -1 * grad_psi_phi[0] + u[1] * -1 * grad_psi_phi[1] + u[2] * -1 * grad_psi_phi[2]) / rho;
            fout[(size_t)k * ne + i] = (1 - omega) * fin[(size_t)k * ne + i] + omega * eqf0 + ff0;
            gout[(size_t)k * ne + i] = (1 - omega) * gin[(size_t)k * ne + i] + omega * eqg0 + fg0;
        }
    }
}

static void hcz3_fields(const clbm_params *p, const double *fin, const double *gin, const uint8_t *flag,
                        double *s0, double *s1, double *s2, double *ux, double *uy, double *uz)
{
    const size_t ne = (size_t)p->nx * p->ny * p->nz;
    hcz3_f F;
    hcz3_fieldsets(p, fin, gin, flag, &F);
#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)ne; ++ii) {
        size_t i = (size_t)ii;
        if (s0) s0[i] = F.phi[i];
        if (s2) s2[i] = F.rho[i];
        int bulk = flag[i] == BULK;
        /* total_P :318-328 */
        if (s1) s1[i] = bulk ? F.Pt[i] - 0.5 * (F.u0[i] * F.gp0[i] + F.u1[i] * F.gp1[i] + F.u2[i] * F.gp2[i]) : 0.0;
        if (ux) ux[i] = bulk ? F.u0[i] : 0.0;
        if (uy) uy[i] = bulk ? F.u1[i] : 0.0;
        if (uz) uz[i] = bulk ? F.u2[i] : 0.0;
    }
}

/* ===========================================================================
 * public oracle entry points (reference layout, whole lattice, nx == nx_global)
 * ======================================================================== */
static int model_Q(int model) { return (model == CLBM_MODEL_SC_D3Q19 || model == CLBM_MODEL_HCZ_D3Q19) ? 19 : 9; }
static int model_sets(int model) { return (model == CLBM_MODEL_HCZ_D2Q9 || model == CLBM_MODEL_HCZ_D3Q19) ? 2 : 1; }

size_t oracle_lattice_size(const clbm_params *p)
{ /* sizeOfLattice: laplace2D.h:93, rayleighTaylor2D.h:101-104, laplace3D.h:118-119 */
    return (size_t)2 * model_sets(p->model) * model_Q(p->model) * ((size_t)p->nx * p->ny * p->nz);
}

int oracle_step(const clbm_params *p, double *lattice, const uint8_t *flag, int *parity, int nsteps, int nthreads)
{
    const size_t ne = (size_t)p->nx * p->ny * p->nz, npop = (size_t)model_Q(p->model) * ne;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    for (int s = 0; s < nsteps; ++s) {
        double *fin = lattice + (size_t)(*parity) * npop, *fout = lattice + (size_t)(1 - *parity) * npop;
        double *gin = fin + 2 * npop, *gout = fout + 2 * npop;
        switch (p->model) {
        case CLBM_MODEL_SC_D2Q9:
            if (p->sc_force == CLBM_SC_FORCE_EXPGUO) scrt_step(p, fin, fout, flag); else sc_step(p, 2, fin, fout, flag);
            break;
        case CLBM_MODEL_SC_D3Q19: sc_step(p, 3, fin, fout, flag); break;
        case CLBM_MODEL_HCZ_D2Q9: hcz2_step(p, fin, fout, gin, gout, flag); break;
        case CLBM_MODEL_HCZ_D3Q19: hcz3_step(p, fin, fout, gin, gout, flag); break;
        default: return -1;
        }
        *parity = 1 - *parity;
    }
    return 0;
}

int oracle_fields(const clbm_params *p, const double *lattice, const uint8_t *flag, int parity,
                  double *s0, double *s1, double *s2, double *ux, double *uy, double *uz)
{
    const size_t ne = (size_t)p->nx * p->ny * p->nz, npop = (size_t)model_Q(p->model) * ne;
    const double *fin = lattice + (size_t)parity * npop, *gin = fin + 2 * npop;
    switch (p->model) {
    case CLBM_MODEL_SC_D2Q9:
        if (p->sc_force == CLBM_SC_FORCE_EXPGUO) scrt_fields(p, fin, flag, s0, s1, ux, uy, uz, NULL, NULL);
        else sc_fields(p, 2, fin, flag, s0, s1, ux, uy, uz);
        break;
    case CLBM_MODEL_SC_D3Q19: sc_fields(p, 3, fin, flag, s0, s1, ux, uy, uz); break;
    case CLBM_MODEL_HCZ_D2Q9:
        hcz2_fields(p, fin, gin, flag, s0, s1, s2, ux, uy);
        if (uz) memset(uz, 0, ne * sizeof(double));
        break;
    case CLBM_MODEL_HCZ_D3Q19: hcz3_fields(p, fin, gin, flag, s0, s1, s2, ux, uy, uz); break;
    default: return -1;
    }
    return 0;
}

/* interaction force of every node (0 at non-bulk nodes): force_ff of the Rayleigh-Taylor variant, `force` otherwise */
int oracle_force(const clbm_params *p, const double *lattice, const uint8_t *flag, int parity, double *fx, double *fy, double *fz)
{
    const size_t ne = (size_t)p->nx * p->ny * p->nz, npop = (size_t)model_Q(p->model) * ne;
    const double *fin = lattice + (size_t)parity * npop;
    if (p->model != CLBM_MODEL_SC_D2Q9 && p->model != CLBM_MODEL_SC_D3Q19) return -1;
    if (p->sc_force == CLBM_SC_FORCE_EXPGUO) {
        if (p->model != CLBM_MODEL_SC_D2Q9) return -1;
        scrt_fields(p, fin, flag, NULL, NULL, NULL, NULL, NULL, fx, fy);
        if (fz) memset(fz, 0, ne * sizeof(double));
        return 0;
    }
    const int D = p->model == CLBM_MODEL_SC_D2Q9 ? 2 : 3;
    const sc_eos e = {p->R, p->TT, p->a};
    double *psi = scratch(0, ne), *rhoa = scratch(1, ne);
    sc_psi_field(p, &e, D, fin, flag, psi, rhoa);
    for (size_t i = 0; i < ne; ++i) {
        int iX = (int)(i / ((size_t)p->ny * p->nz)), rem = (int)(i % ((size_t)p->ny * p->nz)), iY = rem / p->nz, iZ = rem % p->nz;
        double F[3] = {0., 0., 0.};
        if (flag[i] == BULK) sc_force(p, &e, D, psi, flag, rhoa[i], iX, iY, iZ, F);
        if (fx) fx[i] = F[0];
        if (fy) fy[i] = F[1];
        if (fz) fz[i] = F[2];
    }
    return 0;
}

/* ---- initial conditions (iniLattice + inigeom of each case) ---------------- */
int oracle_init_case(const clbm_params *p, int case_id, const double *args, int nargs,
                     double *lattice, uint8_t *flag, int *parity)
{
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    const size_t ne = (size_t)nx * ny * nz;
    const int Q = model_Q(p->model);
    const size_t npop = (size_t)Q * ne;
    memset(lattice, 0, oracle_lattice_size(p) * sizeof(double));
    *parity = 0;
    double *f = lattice, *g = lattice + 2 * npop;
    for (size_t i = 0; i < ne; ++i) {
        int iX = (int)(i / ((size_t)ny * nz)), rem = (int)(i % ((size_t)ny * nz)), iY = rem / nz, iZ = rem % nz;
        int wall = 0;
        switch (case_id) {
        case CLBM_CASE_SC_LAPLACE2D: { /* SC/apps/laplace2D.h:132-145 */
            if (nargs < 3) return -1;
            double cx = (double)nx / 2.0, cy = (double)ny / 2.0, Rdrop = args[2];
            double dx = (double)iX - cx, dy = (double)iY - cy;
            double rho = (dx * dx + dy * dy <= Rdrop * Rdrop) ? args[0] : args[1];
            for (int k = 0; k < 9; ++k) f[(size_t)k * ne + i] = rho * T9[k];
        } break;
        case CLBM_CASE_SC_CONTACT2D: { /* SC/apps/contactAngle2D.h:126-137, 442-455 */
            if (nargs < 3) return -1;
            int x_c = nx / 2, y_c = 5;
            double dx = (double)iX - (double)x_c, dy = (double)iY - (double)y_c;
            double rho = (dx * dx + dy * dy <= args[2] * args[2]) ? args[0] : args[1];
            for (int k = 0; k < 9; ++k) f[(size_t)k * ne + i] = rho * T9[k];
            wall = (iY == 0 || iY == ny - 1);
        } break;
        case CLBM_CASE_SC_LAYERED2D: { /* SC/apps/twoLayeredFlow2D.h:325-346 (iniLattice_layers), :441-454 (inigeom) */
            if (nargs < 4) return -1;
            const double rhol = args[0], rhog = args[1], h_lower = args[2];
            const int w_int = (int)args[3];
            const double H = (double)(ny - 1);
            const double y_low = (h_lower < 0.0 ? 0.0 : (h_lower > 0.5 ? 0.5 : h_lower)) * H;
            const double y_high = H - y_low;
            const double w = (double)(w_int > 1 ? w_int : 1);
            const double yy = (double)iY;
            const double s_bottom = 0.5 * (1.0 - tanh((yy - y_low) / w));
            const double s_top = 0.5 * (1.0 + tanh((yy - y_high) / w));
            double s_liq = s_bottom + s_top;
            s_liq = s_liq < 0.0 ? 0.0 : (s_liq > 1.0 ? 1.0 : s_liq);
            const double s_gas = 1.0 - s_liq;
            const double rho = s_liq * rhog + s_gas * rhol;
            for (int k = 0; k < 9; ++k) f[(size_t)k * ne + i] = rho * T9[k];
            wall = (iY == 0 || iY == ny - 1);
        } break;
        case CLBM_CASE_SC_RT2D: { /* SC/apps/RayleighTaylor2D.h:134-158 (iniLattice), :526-541 (inigeom) */
            if (nargs < 2) return -1;
            const double rhol = args[0], rhog = args[1];
            double x = (double)iX;
            double interface = ((double)ny / 2.0) + ((double)nx) * 0.1 * cos(2.0 * M_PI * x / ((double)(nx - 1)));
            double w = 2.5, y = (double)iY;
            double rho = 0.5 * (rhol + rhog) + 0.5 * (rhol - rhog) * tanh((y - interface) / (2.0 * w));
            for (int k = 0; k < 9; ++k) f[(size_t)k * ne + i] = rho * T9[k];
            wall = (iY == 0 || iY == ny - 1);
        } break;
        case CLBM_CASE_SC_DROPLET3D: /* composed: contactAngle2D geometry extruded to 3-D, sphere centre (nx/2, yc, nz/2) */
        case CLBM_CASE_SC_DROPLET3D_PER: {
            if (nargs < 3) return -1;
            int per = case_id == CLBM_CASE_SC_DROPLET3D_PER;
            double yc = per ? (double)(ny / 2) : (nargs > 3 ? args[3] : 5.0);
            double dx = (double)iX - (double)(nx / 2), dy = (double)iY - yc, dz = (double)iZ - (double)(nz / 2);
            double rho = (dx * dx + dy * dy + dz * dz <= args[2] * args[2]) ? args[0] : args[1];
            for (int k = 0; k < 19; ++k) f[(size_t)k * ne + i] = rho * T19[k];
            wall = !per && (iY == 0 || iY == ny - 1);
        } break;
        case CLBM_CASE_HCZ_RT2D: { /* PF/apps/rayleighTaylor2D.h:155-193, 802-820 */
            double x = (double)iX;
            double interface = ((double)ny / 2.0) + ((double)nx) * 0.1 * cos(2.0 * M_PI * x / ((double)(nx - 1)));
            double w = 1.25, y = (double)iY;
            double phi = 0.5 * (p->phi_l + p->phi_g) + 0.5 * (p->phi_l - p->phi_g) * tanh((y - interface) / (2.0 * w));
            double rho = p->rho_g + ((phi - p->phi_g) / (p->phi_l - p->phi_g)) * (p->rho_l - p->rho_g);
            double rt_rho = p->b * rho / 4.0;
            double p_rho = (rho / 3.0) * (1.0 + rt_rho + rt_rho * rt_rho - rt_rho * rt_rho * rt_rho) / pow(1.0 - rt_rho, 3) - p->a * rho * rho;
            for (int k = 0; k < 9; ++k) { f[(size_t)k * ne + i] = phi * T9[k]; g[(size_t)k * ne + i] = p_rho * T9[k]; }
            wall = (iY == 0 || iY == ny - 1);
        } break;
        case CLBM_CASE_HCZ_LAYERED2D: { /* PF/apps/twoLayeredFlow2D.h:148-196 (iniLattice_layers: BOTH buffers), :737-757 (inigeom) */
            if (nargs < 2) return -1;
            const double h_lower = args[0];
            const int w_int = (int)args[1];
            const double H = (double)(ny - 1);
            const double y_low = (h_lower < 0.0 ? 0.0 : (h_lower > 0.5 ? 0.5 : h_lower)) * H, y_high = H - y_low;
            const double w = (double)(w_int > 1 ? w_int : 1), yy = (double)iY;
            const double s_bottom = 0.5 * (1.0 - tanh((yy - y_low) / w));
            const double s_top = 0.5 * (1.0 + tanh((yy - y_high) / w));
            double s_liq = s_bottom + s_top;
            s_liq = s_liq < 0.0 ? 0.0 : (s_liq > 1.0 ? 1.0 : s_liq);
            const double s_gas = 1.0 - s_liq;
            const double phi = s_liq * p->phi_g + s_gas * p->phi_l;
            const double rho = s_liq * p->rho_g + s_gas * p->rho_l;
            const double rt = p->b * rho / 4.0;
            const double denom = pow(1.0 - rt, 3);
            const double p_rho = (rho / 3.0) * (1.0 + rt + rt * rt - rt * rt * rt) / denom - p->a * rho * rho;
            for (int k = 0; k < 9; ++k) {
                f[(size_t)k * ne + i] = phi * T9[k];
                g[(size_t)k * ne + i] = p_rho * T9[k];
                f[npop + (size_t)k * ne + i] = phi * T9[k];      /* fout = fin, gout = gin (:189-190); inigeom only */
                g[npop + (size_t)k * ne + i] = p_rho * T9[k];    /* zeroes the parity-0 buffer of the wall nodes     */
            }
            wall = (iY == 0 || iY == ny - 1);
        } break;
        case CLBM_CASE_HCZ_LAPLACE3D: { /* PF/apps/laplace3D.h:170-213 */
            double xc = (double)nx / 2.0, yc = (double)ny / 2.0, zc = (double)nz / 2.0, R = 0.25 * nx;
            const double xi = 1.0;
            double dx = (double)iX - xc, dy = (double)iY - yc, dz = (double)iZ - zc;
            double delta = sqrt(dx * dx + dy * dy + dz * dz) - R;
            double rt_l = p->b * p->phi_l / 4.0;
            double pth_l = (p->phi_l / 3.0) * (1 + rt_l + rt_l * rt_l - rt_l * rt_l * rt_l) / pow(1 - rt_l, 3) - p->a * p->phi_l * p->phi_l;
            double rt_g = p->b * p->phi_g / 4.0;
            double pth_g = (p->phi_g / 3.0) * (1 + rt_g + rt_g * rt_g - rt_g * rt_g * rt_g) / pow(1 - rt_g, 3) - p->a * p->phi_g * p->phi_g;
            double w = 0.5 * (1.0 - tanh(delta / xi));
            double phi = p->phi_g + w * (p->phi_l - p->phi_g);
            double pth = pth_g + w * (pth_l - pth_g);
            for (int k = 0; k < 19; ++k) { f[(size_t)k * ne + i] = phi * T19[k]; g[(size_t)k * ne + i] = pth * T19[k]; }
        } break;
        default: return -1;
        }
        flag[i] = wall ? BB : BULK;
        if (wall) {
            for (int k = 0; k < Q; ++k) {
                f[(size_t)k * ne + i] = 0.0;
                if (model_sets(p->model) == 2) g[(size_t)k * ne + i] = 0.0;
            }
        }
    }
    return 0;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
