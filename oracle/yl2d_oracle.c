/*
 * yl2d_oracle.c -- CPU ORACLE (test infrastructure, NOT a product path) of "Abbashub LBM/apps/Young_Laplace2D.h" (AB/
 * below): conservative phase-field LBM (Fakhari et al. 2017) with a velocity-based hydrodynamic population, D2Q9 BGK,
 * fully periodic -- the problem the AB reference build runs by default (AB/apps/COOLBM.cpp:99).
 *
 * Plain-C restatement of LBM_Young_Laplace2D: iniCell (AB:141-169), collide_stream_at (:217-290), update_fields
 * (:297-370) -- same expressions, same association, same loop order, -ffp-contract=off.  PINNED bit-for-bit against
 * the untouched header (oracle/_ref/ref_yl2d, fixtures tests/golden/yl2d_*.npz, tests/test_yl2d_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const int CK[9][2] = {{-1, 0}, {0, -1}, {-1, -1}, {-1, 1}, {0, 0}, {1, 0}, {0, 1}, {1, 1}, {1, -1}};
static const int OPP[9] = {5, 6, 7, 8, 4, 0, 1, 2, 3};
static const double TK[9] = {1. / 9., 1. / 9., 1. / 36., 1. / 36., 4. / 9., 1. / 9., 1. / 9., 1. / 36., 1. / 36.};

typedef struct yl2d {
    int nx, ny, parity;
    size_t ne;
    double *lattice;   /* [h_in | h_out | g_in | g_out], 4*9*ne */
    double Rhol, Rhoh, Sigma, W, M, tau, s8, Beta, kappa, dRho3;
    double *C, *P, *Rho, *Ux, *Uy, *mu, *DcDx, *DcDy, *ni, *nj;
} yl2d;

#define FIN(s, i, k) ((s)->lattice[(size_t)(s)->parity * 9 * (s)->ne + (size_t)(k) * (s)->ne + (i)])
#define FOUT(s, i, k) ((s)->lattice[(size_t)(1 - (s)->parity) * 9 * (s)->ne + (size_t)(k) * (s)->ne + (i)])
#define GIN(s, i, k) ((s)->lattice[2 * 9 * (s)->ne + (size_t)(s)->parity * 9 * (s)->ne + (size_t)(k) * (s)->ne + (i)])
#define GOUT(s, i, k) ((s)->lattice[2 * 9 * (s)->ne + (size_t)(1 - (s)->parity) * 9 * (s)->ne + (size_t)(k) * (s)->ne + (i)])
static int wrap(int v, int n) { v %= n; return v < 0 ? v + n : v; }
static size_t at(const yl2d *s, int X, int Y) { return (size_t)Y + (size_t)s->ny * (size_t)X; }

/* AB:174-180 */
static void gawa(double U, double V, double out[9])
{
    double U2 = U * U + V * V;
    for (int k = 0; k < 9; ++k) {
        double eU = CK[k][0] * U + CK[k][1] * V;
        out[k] = TK[k] * (3.0 * eU + 4.5 * eU * eU - 1.5 * U2);
    }
}
/* AB:183-201 */
static void viscous_force(const yl2d *s, double dcdx, double dcdy, const double gneq[9], double *FmX, double *FmY)
{
    double sxx = 0.0, sxy = 0.0, syy = 0.0;
    for (int k = 0; k < 9; ++k) {
        if (k == 4) continue;
        sxx += gneq[k] * CK[k][0] * CK[k][0];
        sxy += gneq[k] * CK[k][0] * CK[k][1];
        syy += gneq[k] * CK[k][1] * CK[k][1];
    }
    double fac = (0.5 - s->tau) / s->tau;
    double dR = (s->Rhoh - s->Rhol);
    *FmX = fac * (sxx * dcdx + sxy * dcdy) * dR;
    *FmY = fac * (sxy * dcdx + syy * dcdy) * dR;
}

/* AB:141-169 */
static void ini_cell(yl2d *s, int i)
{
    int X = i / s->ny, Y = i % s->ny;
    double xc = (double)s->nx / 2.0 - 0.5, yc = (double)s->ny / 2.0 - 0.5;
    double R0 = (double)s->nx / 8.0;
    double r = sqrt((X - xc) * (X - xc) + (Y - yc) * (Y - yc));
    double phi = 0.5 - 0.5 * tanh(2.0 * (R0 - r) / s->W);
    for (int k = 0; k < 9; ++k) { FIN(s, i, k) = phi * TK[k]; GIN(s, i, k) = 0.0; }
    s->C[i] = phi;
    s->Rho[i] = s->Rhol + phi * (s->Rhoh - s->Rhol);
    s->Ux[i] = s->Uy[i] = 0.0;
    s->P[i] = 0.0;
    double prho = (s->Rho[i] + 1e-12) / 3.0;
    double corr = (phi * s->Sigma / R0) / prho;
    s->P[i] -= corr;
    for (int k = 0; k < 9; ++k) GIN(s, i, k) = s->P[i] * TK[k];
}

/* AB:297-370 */
static void update_fields(yl2d *s)
{
    const int nx = s->nx, ny = s->ny;
    for (int X = 0; X < nx; ++X)
        for (int Y = 0; Y < ny; ++Y) {
            size_t i = at(s, X, Y);
            double phi = 0.0;
            for (int k = 0; k < 9; ++k) phi += FIN(s, i, k);
            s->C[i] = phi;
            s->Rho[i] = s->Rhol + phi * (s->Rhoh - s->Rhol);
        }
#define CAT(X_, Y_) s->C[at(s, wrap((X_), nx), wrap((Y_), ny))]
    for (int X = 0; X < nx; ++X)
        for (int Y = 0; Y < ny; ++Y) {
            size_t i = at(s, X, Y);
            double cE = CAT(X + 1, Y), cW = CAT(X - 1, Y), cN = CAT(X, Y + 1), cS = CAT(X, Y - 1);
            double cNE = CAT(X + 1, Y + 1), cNW = CAT(X - 1, Y + 1), cSE = CAT(X + 1, Y - 1), cSW = CAT(X - 1, Y - 1);
            s->DcDx[i] = (cE - cW) / 3.0 + (cSE + cNE - cSW - cNW) / 12.0;
            s->DcDy[i] = (cN - cS) / 3.0 + (cNW + cNE - cSW - cSE) / 12.0;
        }
    for (int X = 0; X < nx; ++X)
        for (int Y = 0; Y < ny; ++Y) {
            size_t i = at(s, X, Y);
            double cC = CAT(X, Y);
            double cE = CAT(X + 1, Y), cW = CAT(X - 1, Y), cN = CAT(X, Y + 1), cS = CAT(X, Y - 1);
            double cNE = CAT(X + 1, Y + 1), cNW = CAT(X - 1, Y + 1), cSE = CAT(X + 1, Y - 1), cSW = CAT(X - 1, Y - 1);
            double D2C = (cSW + cSE + cNW + cNE + 4.0 * (cS + cW + cE + cN) - 20.0 * cC) / 6.0;
            s->mu[i] = 4.0 * s->Beta * cC * (cC - 1.0) * (cC - 0.5) - s->kappa * D2C;
        }
#undef CAT
    for (size_t i = 0; i < s->ne; ++i) {
        double g2 = s->DcDx[i] * s->DcDx[i] + s->DcDy[i] * s->DcDy[i] + 1e-32;
        double inv = 1.0 / sqrt(g2);
        s->ni[i] = s->DcDx[i] * inv;
        s->nj[i] = s->DcDy[i] * inv;
    }
    for (int X = 0; X < nx; ++X)
        for (int Y = 0; Y < ny; ++Y) {
            size_t i = at(s, X, Y);
            double pstar = 0.0;
            for (int k = 0; k < 9; ++k) pstar += GIN(s, i, k);
            s->P[i] = pstar;
            double FpX = -s->P[i] * s->dRho3 * s->DcDx[i];
            double FpY = -s->P[i] * s->dRho3 * s->DcDy[i];
            double GaWa[9], geqk[9], gneq[9];
            gawa(s->Ux[i], s->Uy[i], GaWa);
            for (int k = 0; k < 9; ++k) geqk[k] = s->P[i] * TK[k] + GaWa[k];
            for (int k = 0; k < 9; ++k) gneq[k] = GIN(s, i, k) - geqk[k];
            double FmX, FmY;
            viscous_force(s, s->DcDx[i], s->DcDy[i], gneq, &FmX, &FmY);
            double Fx = s->mu[i] * s->DcDx[i] + FpX + FmX;
            double Fy = s->mu[i] * s->DcDy[i] + FpY + FmY;
            double mx = 0.0, my = 0.0;
            for (int k = 0; k < 9; ++k) { mx += GIN(s, i, k) * CK[k][0]; my += GIN(s, i, k) * CK[k][1]; }
            s->Ux[i] = mx + 0.5 * Fx / (s->Rho[i] + 1e-30);
            s->Uy[i] = my + 0.5 * Fy / (s->Rho[i] + 1e-30);
        }
}

/* AB:217-290 */
static void collide_stream_at(yl2d *s, int i)
{
    const int X = i / s->ny, Y = i % s->ny;
    double Cc = s->C[i], Rhoc = s->Rho[i], Pn = s->P[i], U = s->Ux[i], V = s->Uy[i];
    double dCx = s->DcDx[i], dCy = s->DcDy[i], muc = s->mu[i], ni_ = s->ni[i], nj_ = s->nj[i];
    double GaWa[9], GammaK[9];
    gawa(U, V, GaWa);
    for (int k = 0; k < 9; ++k) GammaK[k] = TK[k] + GaWa[k];
    double hlp_h[9], heqk[9];
    double shape = (1.0 - 4.0 * (Cc - 0.5) * (Cc - 0.5)) / s->W;
    for (int k = 0; k < 9; ++k) {
        double proj = CK[k][0] * ni_ + CK[k][1] * nj_;
        double eFh = shape * proj;
        hlp_h[k] = TK[k] * eFh;
        heqk[k] = Cc * GammaK[k] - 0.5 * hlp_h[k];
    }
    double wc = 1.0 / (0.5 + 3.0 * s->M);
    double FpX = -Pn * s->dRho3 * dCx, FpY = -Pn * s->dRho3 * dCy;
    double geqk[9], gneq[9];
    for (int k = 0; k < 9; ++k) geqk[k] = Pn * TK[k] + GaWa[k];
    for (int k = 0; k < 9; ++k) gneq[k] = GIN(s, i, k) - geqk[k];
    double FmX, FmY;
    viscous_force(s, dCx, dCy, gneq, &FmX, &FmY);
    double Fx = muc * dCx + FpX + FmX, Fy = muc * dCy + FpY + FmY;
    double hlp_g[9], geq_corr[9];
    for (int k = 0; k < 9; ++k) {
        double eF = CK[k][0] * Fx + CK[k][1] * Fy;
        hlp_g[k] = 3.0 * TK[k] * eF / (Rhoc + 1e-30);
        geq_corr[k] = geqk[k] - 0.5 * hlp_g[k];
    }
    for (int k = 0; k < 9; ++k) {
        double hk = (1.0 - wc) * FIN(s, i, k) + wc * heqk[k] + hlp_h[k];
        double gk = (1.0 - s->s8) * GIN(s, i, k) + s->s8 * geq_corr[k] + hlp_g[k];
        size_t nb = (k == 4) ? (size_t)i : at(s, wrap(X + CK[k][0], s->nx), wrap(Y + CK[k][1], s->ny));
        FOUT(s, nb, k) = hk;
        GOUT(s, nb, k) = gk;
    }
    (void)OPP;
}

/* ---- public API (ctypes) ---- */
void yl2d_destroy(yl2d *s)
{
    if (!s) return;
    free(s->lattice); free(s->C); free(s->P); free(s->Rho); free(s->Ux); free(s->Uy); free(s->mu);
    free(s->DcDx); free(s->DcDy); free(s->ni); free(s->nj);
    free(s);
}
/* driver set-up AB:499-526: parameters, iniCell on every node, update_fields */
yl2d *yl2d_create(int nx, int ny, double Sigma, double W, double M, double RhoL, double RhoH, double tau)
{
    yl2d *s = (yl2d *)calloc(1, sizeof(yl2d));
    s->nx = nx; s->ny = ny; s->ne = (size_t)nx * ny;
    s->lattice = (double *)calloc(4 * 9 * s->ne, sizeof(double));
    double **f[10] = {&s->C, &s->P, &s->Rho, &s->Ux, &s->Uy, &s->mu, &s->DcDx, &s->DcDy, &s->ni, &s->nj};
    for (int j = 0; j < 10; ++j) *f[j] = (double *)calloc(s->ne, sizeof(double));
    s->Sigma = Sigma; s->W = W; s->M = M; s->Rhol = RhoL; s->Rhoh = RhoH; s->tau = tau; s->s8 = 1.0 / tau;
    s->Beta = 12.0 * s->Sigma / s->W;
    s->kappa = 1.5 * s->Sigma * s->W;
    s->dRho3 = (s->Rhoh - s->Rhol) / 3.0;
    for (size_t i = 0; i < s->ne; ++i) ini_cell(s, (int)i);
    update_fields(s);
    return s;
}
void yl2d_step(yl2d *s, int n)
{
    for (int it = 0; it < n; ++it) {
        for (size_t i = 0; i < s->ne; ++i) collide_stream_at(s, (int)i);
        s->parity = 1 - s->parity;
        update_fields(s);
    }
}
int yl2d_parity(const yl2d *s) { return s->parity; }
const double *yl2d_lattice(const yl2d *s) { return s->lattice; }
void yl2d_get(const yl2d *s, double *C, double *P, double *Rho, double *Ux, double *Uy)
{
    const size_t n = s->ne * sizeof(double);
    if (C) memcpy(C, s->C, n);
    if (P) memcpy(P, s->P, n);
    if (Rho) memcpy(Rho, s->Rho, n);
    if (Ux) memcpy(Ux, s->Ux, n);
    if (Uy) memcpy(Uy, s->Uy, n);
}
