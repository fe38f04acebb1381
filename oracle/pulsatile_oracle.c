/*
 * pulsatile_oracle.c -- CPU ORACLE (test infrastructure, NOT a product path) of the compliant-vessel case
 * "Abbashub LBM/apps/PulsatileBloodFlow2D.h" (AB/ below): pressure-based D2Q9 MRT, Bouzidi curved moving
 * walls, pull streaming, Zou/He pressure inlet/outlet, pressure-driven wall motion with fresh-node filling.
 *
 * Plain-C restatement of the reference time step in the reference's own order of operations
 *   collide -> Bouzidi -> pull stream -> Zou/He in/out -> macroscopic -> move walls -> parity flip
 * (AB/apps/PulsatileBloodFlow2D.h:764-790), same expressions, same association, same loop order, built with
 * -ffp-contract=off, so every double is meant to be bit-identical to the reference.  The reference quirks of
 * SURVEY.md Appendix B.1-B.4 are kept: the parity flip after the in-place pull, the k-ordered arrays fed to the
 * I-ordered moment transform, wrapY == identity (flat-index spill into the neighbouring column).
 *
 * PINNED: tests/test_pulsatile_oracle.py checks it (a) bit-for-bit against binary dumps of the untouched
 * reference header (oracle/_ref/ref_pulsatile, fixtures in tests/golden/) and (b) byte-for-byte against the
 * SHA-256 of all 103 VTK files the reference ships in
 * "Abbashub LBM/out_single-phase fluid flow through a compliant vessel/" (N = 64).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* AB:29-38 (k ordering) and :41-49 (Abbas "I" ordering 0 rest, 1 E, 2 N, 3 W, 4 S, 5 NE, 6 NW, 7 SW, 8 SE) */
static const int CK[9][2] = {{-1, 0}, {0, -1}, {-1, -1}, {-1, 1}, {0, 0}, {1, 0}, {0, 1}, {1, 1}, {1, -1}};
static const double TK[9] = {1. / 9., 1. / 9., 1. / 36., 1. / 36., 4. / 9., 1. / 9., 1. / 9., 1. / 36., 1. / 36.};
static const int EXI[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
static const int EYI[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
static const int JBI[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
static const int KFROMI[9] = {4, 5, 6, 0, 1, 7, 3, 2, 8};

typedef struct { int X, Y; double Delta[8]; } border_node;

typedef struct pulsatile {
    int nx, ny;
    size_t nelem, npop;
    double *lattice; /* 2*9*nelem doubles (+1 zeroed pad: the reference reads one element past the end, B.3) */
    uint8_t *flag;   /* 0 bounce_back, 1 bulk */
    int parity;
    double Rho0, tau, s8, s5, S[9];
    int deformable, is_severed;
    double alpha, p0_in, p0_out, p_tissue, p_oscillatory, Delta_p, omega;
    int t_beat, t_propagation, t_start, t_sever;
    double *P, *Ux, *Uy, *yr1, *yr2, *y1new, *y2new, *Vw1, *Vw2, *Fobj, *Fold;
    border_node *B1, *B2;
    int Nb1, Nb2, cap1, cap2;
    int FreshNodes, KilledNodes;
    int t_iter;
} pulsatile;

#define XY(p, x, y) ((y) + (p)->ny * (x))
#define GIN(p, i, k) ((p)->lattice[(size_t)(p)->parity * (p)->npop + (size_t)(k) * (p)->nelem + (i)])
#define GOUT(p, i, k) ((p)->lattice[(size_t)(1 - (p)->parity) * (p)->npop + (size_t)(k) * (p)->nelem + (i)])
#define FO(p, Xp, Yp) ((p)->Fobj[(Xp) * ((p)->ny + 2) + (Yp)])
static int X0c(const pulsatile *p) { (void)p; return 0; }
static int Y0c(const pulsatile *p) { return (p->ny - 1) / 2; }
static int solid(const pulsatile *p, int X, int Y) { return p->flag[XY(p, X, Y)] == 0; }

/* AB:501-507 */
static void equilibrium_g(const pulsatile *p, double P_, double U, double V, double geq[9])
{
    double U2 = U * U + V * V;
    for (int k = 0; k < 9; ++k) {
        double eU = CK[k][0] * U + CK[k][1] * V;
        geq[k] = TK[k] * (P_ + p->Rho0 / 3.0 * (eU * (3.0 + 4.5 * eU) - 1.5 * U2));
    }
}

/* AB:509-531 */
static void convert(const double IN[9], double OUT[9])
{
    OUT[0] = IN[0] + IN[1] + IN[2] + IN[3] + IN[4] + IN[5] + IN[6] + IN[7] + IN[8];
    OUT[1] = -IN[1] - IN[2] - IN[3] - IN[4] + 2 * (IN[5] + IN[6] + IN[7] + IN[8]) - 4 * IN[0];
    OUT[2] = (IN[5] + IN[6] + IN[7] + IN[8]) - 2 * (IN[1] + IN[2] + IN[3] + IN[4]) + 4 * IN[0];
    OUT[3] = IN[1] - IN[3] + IN[5] - IN[6] - IN[7] + IN[8];
    OUT[4] = IN[5] - IN[6] - IN[7] + IN[8] - 2 * (IN[1] - IN[3]);
    OUT[5] = IN[2] - IN[4] + IN[5] + IN[6] - IN[7] - IN[8];
    OUT[6] = IN[5] + IN[6] - IN[7] - IN[8] - 2 * (IN[2] - IN[4]);
    OUT[7] = IN[1] - IN[2] + IN[3] - IN[4];
    OUT[8] = IN[5] - IN[6] + IN[7] - IN[8];
}
static void reconvert(const double IN[9], double OUT[9])
{
    double C0 = IN[0] / 9.0, C7 = IN[7] / 4.0, C8 = IN[8] / 4.0;
    OUT[0] = C0 - (IN[1] - IN[2]) / 9.0;
    OUT[1] = C0 - (IN[1] + 2 * IN[2]) / 36.0 + (IN[3] - IN[4]) / 6.0 + C7;
    OUT[2] = C0 - (IN[1] + 2 * IN[2]) / 36.0 + (IN[5] - IN[6]) / 6.0 - C7;
    OUT[3] = C0 - (IN[1] + 2 * IN[2]) / 36.0 - (IN[3] - IN[4]) / 6.0 + C7;
    OUT[4] = C0 - (IN[1] + 2 * IN[2]) / 36.0 - (IN[5] - IN[6]) / 6.0 - C7;
    OUT[5] = C0 + (IN[2] + 2 * IN[1]) / 36.0 + (IN[3] + IN[5]) / 6.0 + (IN[4] + IN[6]) / 12.0 + C8;
    OUT[6] = C0 + (IN[2] + 2 * IN[1]) / 36.0 - (IN[3] - IN[5]) / 6.0 - (IN[4] - IN[6]) / 12.0 - C8;
    OUT[7] = C0 + (IN[2] + 2 * IN[1]) / 36.0 - (IN[3] + IN[5]) / 6.0 - (IN[4] + IN[6]) / 12.0 + C8;
    OUT[8] = C0 + (IN[2] + 2 * IN[1]) / 36.0 + (IN[3] - IN[5]) / 6.0 + (IN[4] - IN[6]) / 12.0 - C8;
}

/* AB:533-541 (the only part the reference runs in parallel) */
static void mrt_collision(pulsatile *p, int X, int Y)
{
    int i = XY(p, X, Y);
    if (solid(p, X, Y)) return;
    double geqv[9], tmp[9], m[9], dpost[9];
    equilibrium_g(p, p->P[i], p->Ux[i], p->Uy[i], geqv);
    for (int k = 0; k < 9; ++k) tmp[k] = GIN(p, i, k) - geqv[k];
    convert(tmp, m);
    for (int q = 0; q < 9; ++q) m[q] *= p->S[q];
    reconvert(m, dpost);
    for (int k = 0; k < 9; ++k) GOUT(p, i, k) = GIN(p, i, k) - dpost[k];
}

/* AB:147-168 */
static void setup_parameters(pulsatile *p)
{
    if (p->t_beat <= 0) p->t_beat = p->nx > 1 ? p->nx : 1;
    p->omega = 2.0 * 3.141592653589793 / (double)p->t_beat;
    if (p->p0_in == 0.0 && p->p0_out == 0.0) { p->p0_in = 0.20; p->p0_out = 0.19; }
    if (p->is_severed) { p->p0_in = 0.02; p->p0_out = 0.00; }
    p->p_tissue = p->p0_in;
    p->p_oscillatory = (p->p0_in - p->p0_out);
    if (p->is_severed) p->p_oscillatory *= 0.1;
    p->Delta_p = p->p0_out - p->p0_in;
    p->t_propagation = (int)((p->nx - 1.) * sqrt(3.) - 1) * 1;
    p->t_start = 2 * p->t_propagation;
    p->t_sever = 0;
}

/* AB:172-189; returns -1 for "Initial wall location out of bounds." */
static int theoretical_wall_and_pressure(pulsatile *p)
{
    const int ny = p->ny, nx = p->nx;
    double c = Y0c(p) + 0.5;
    double yr1_in = c - (p->p0_in - p->p_tissue) / p->alpha;
    double yr2_in = c + (p->p0_in - p->p_tissue) / p->alpha;
    double yr1_out = c - (p->p0_out - p->p_tissue) / p->alpha;
    double yr2_out = c + (p->p0_out - p->p_tissue) / p->alpha;
    if (yr1_in < 1 || yr2_in > ny - 2 || yr1_out < 1 || yr2_out > ny - 2) return -1;
    double R0 = (yr2_in - yr1_in) / 2.0, RL = (yr2_out - yr1_out) / 2.0;
    for (int X = 0; X < nx; ++X) {
        double Rx4 = (pow(RL, 4) - pow(R0, 4)) * ((double)X / (double)(nx - 1)) + pow(R0, 4);
        double Rx = pow(Rx4, 0.25);
        p->yr1[X] = c - Rx;
        p->yr2[X] = c + Rx;
        for (int Y = 0; Y < ny; ++Y) p->P[XY(p, X, Y)] = (p->yr2[X] - (ny - 1 - 0.5)) * p->alpha + p->p_tissue;
    }
    return 0;
}

/* AB:275-285 */
static void init_fobj(pulsatile *p)
{
    const int NX = p->nx, NY = p->ny;
    const double c = Y0c(p) + 0.5;
    for (int X = 0; X < NX; ++X) {
        for (int Y = -1; Y <= Y0c(p); ++Y) FO(p, X + 1, Y + 1) = (p->yr1[X] - c) / (Y - c);
        for (int Y = Y0c(p) + 1; Y < NY + 1; ++Y) FO(p, X + 1, Y + 1) = (p->yr2[X] - c) / (Y - c);
    }
    for (int Y = 0; Y < NY + 2; ++Y) {
        FO(p, 0, Y) = 2.0 * FO(p, 1, Y) - FO(p, 2, Y);
        FO(p, NX + 1, Y) = 2.0 * FO(p, NX, Y) - FO(p, NX - 1, Y);
    }
    for (int X = 0; X < NX; ++X)
        for (int Y = 0; Y < NY; ++Y) p->flag[XY(p, X, Y)] = (FO(p, X + 1, Y + 1) < 1.0 ? 0 : 1);
}

/* AB:288-290 */
static void find_delta(int mA, double mB, double Y1, double *Delta)
{
    *Delta = 1.0 - fabs(Y1 / (mA - mB));
    if (*Delta < 0) *Delta = 0;
}
static void push_node(border_node **B, int *n, int *cap, int X, int Y, const double D[8])
{
    if (*n == *cap) { *cap = *cap ? 2 * *cap : 256; *B = (border_node *)realloc(*B, (size_t)*cap * sizeof(border_node)); }
    (*B)[*n].X = X;
    (*B)[*n].Y = Y;
    memcpy((*B)[*n].Delta, D, 8 * sizeof(double));
    ++*n;
}
static void d_reset(double D[8]) { for (int i = 0; i < 8; ++i) D[i] = 2; }

/* AB:294-337 */
static void update_boundary_bottom(pulsatile *p)
{
    const double *yr1 = p->yr1;
    const int nx = p->nx;
    p->Nb1 = 0;
    double D[8];
    int X = 0, Y = (int)floor(yr1[X]);
    if (FO(p, X + 1, Y + 1) >= 1) Y = Y - 1;
#define PUSH1(Xn, Yn) push_node(&p->B1, &p->Nb1, &p->cap1, (Xn), (Yn), D)
    d_reset(D);
    if (FO(p, X + 2, Y + 1) >= 1) find_delta(0, yr1[X + 1] - yr1[X], yr1[X] - Y, &D[0]);
    D[1] = 1 - (yr1[X] - Y);
    if (FO(p, X + 2, Y + 2) >= 1) find_delta(1, yr1[X + 1] - yr1[X], yr1[X] - Y, &D[4]);
    PUSH1(X, Y);
    for (X = 1; X < nx - 1; ++X) {
        int Yx = (int)floor(yr1[X]);
        if (FO(p, X + 1, Yx + 1) >= 1) Yx = Yx - 1;
        if (Yx != Y) {
            d_reset(D);
            if (Yx > Y) { find_delta(-1, yr1[X] - yr1[X - 1], yr1[X] - Y, &D[5]); PUSH1(X, Y); }
            else { find_delta(1, yr1[X] - yr1[X - 1], yr1[X - 1] - Yx, &D[4]); PUSH1(X - 1, Yx); }
        }
        d_reset(D);
        if (FO(p, X + 2, Yx + 1) >= 1) find_delta(0, yr1[X + 1] - yr1[X], yr1[X] - Yx, &D[0]);
        D[1] = 1 - (yr1[X] - Yx);
        if (FO(p, X, Yx + 1) >= 1) find_delta(0, yr1[X] - yr1[X - 1], yr1[X] - Yx, &D[2]);
        if (FO(p, X + 2, Yx + 2) >= 1) find_delta(1, yr1[X + 1] - yr1[X], yr1[X] - Yx, &D[4]);
        if (FO(p, X, Yx + 2) >= 1) find_delta(-1, yr1[X] - yr1[X - 1], yr1[X] - Yx, &D[5]);
        PUSH1(X, Yx);
        Y = Yx;
    }
    X = nx - 1;
    int Yx = (int)floor(yr1[X]);
    if (FO(p, X + 1, Yx + 1) >= 1) Yx = Yx - 1;
    if (Yx != Y) {
        d_reset(D);
        if (Yx > Y) { find_delta(-1, yr1[X] - yr1[X - 1], yr1[X] - Y, &D[5]); PUSH1(X, Y); }
        else { find_delta(1, yr1[X] - yr1[X - 1], yr1[X - 1] - Yx, &D[4]); PUSH1(X - 1, Yx); }
    }
    d_reset(D);
    D[1] = 1 - (yr1[X] - Yx);
    if (FO(p, X, Yx + 1) >= 1) find_delta(0, yr1[X] - yr1[X - 1], yr1[X] - Yx, &D[2]);
    if (FO(p, X, Yx + 2) >= 1) find_delta(-1, yr1[X] - yr1[X - 1], yr1[X] - Yx, &D[5]);
    PUSH1(X, Yx);
#undef PUSH1
}

/* AB:339-382 */
static void update_boundary_top(pulsatile *p)
{
    const double *yr2 = p->yr2;
    const int nx = p->nx;
    p->Nb2 = 0;
    double D[8];
    int X = 0, Y = (int)ceil(yr2[X]);
    if (FO(p, X + 1, Y + 1) >= 1) Y = Y + 1;
#define PUSH2(Xn, Yn) push_node(&p->B2, &p->Nb2, &p->cap2, (Xn), (Yn), D)
    d_reset(D);
    if (FO(p, X + 2, Y + 1) >= 1) find_delta(0, yr2[X + 1] - yr2[X], yr2[X] - Y, &D[0]);
    D[3] = 1 - (Y - yr2[X]);
    if (FO(p, X + 2, Y) >= 1) find_delta(-1, yr2[X + 1] - yr2[X], yr2[X] - Y, &D[7]);
    PUSH2(X, Y);
    int Yprev = Y;
    for (X = 1; X < nx - 1; ++X) {
        int Yx = (int)ceil(yr2[X]);
        if (FO(p, X + 1, Yx + 1) >= 1) Yx = Yx + 1;
        if (Yx != Yprev) {
            d_reset(D);
            if (Yx > Yprev) { find_delta(-1, yr2[X] - yr2[X - 1], yr2[X - 1] - Yx, &D[7]); PUSH2(X - 1, Yx); }
            else { find_delta(1, yr2[X] - yr2[X - 1], yr2[X] - Yprev, &D[6]); PUSH2(X, Yprev); }
        }
        d_reset(D);
        if (FO(p, X + 2, Yx + 1) >= 1) find_delta(0, yr2[X + 1] - yr2[X], yr2[X] - Yx, &D[0]);
        if (FO(p, X, Yx + 1) >= 1) find_delta(0, yr2[X] - yr2[X - 1], yr2[X] - Yx, &D[2]);
        D[3] = 1 - (Yx - yr2[X]);
        if (FO(p, X, Yx) >= 1) find_delta(1, yr2[X] - yr2[X - 1], yr2[X] - Yx, &D[6]);
        if (FO(p, X + 2, Yx) >= 1) find_delta(-1, yr2[X + 1] - yr2[X], yr2[X] - Yx, &D[7]);
        PUSH2(X, Yx);
        Yprev = Yx;
    }
    X = nx - 1;
    int Yx = (int)ceil(yr2[X]);
    if (FO(p, X + 1, Yx + 1) >= 1) Yx = Yx + 1;
    if (Yx != Yprev) {
        d_reset(D);
        if (Yx > Yprev) { find_delta(-1, yr2[X] - yr2[X - 1], yr2[X - 1] - Yx, &D[7]); PUSH2(X - 1, Yx); }
        else { find_delta(1, yr2[X] - yr2[X - 1], yr2[X] - Yprev, &D[6]); PUSH2(X, Yprev); }
    }
    d_reset(D);
    if (FO(p, X, Yx + 1) >= 1) find_delta(0, yr2[X] - yr2[X - 1], yr2[X] - Yx, &D[2]);
    D[3] = 1 - (Yx - yr2[X]);
    if (FO(p, X, Yx) >= 1) find_delta(1, yr2[X] - yr2[X - 1], yr2[X] - Yx, &D[6]);
    PUSH2(X, Yx);
#undef PUSH2
}

/* AB:191-214 */
static void initialize_P_U_g(pulsatile *p)
{
    const int nx = p->nx, ny = p->ny;
    for (size_t i = 0; i < p->nelem; ++i) p->Ux[i] = p->Uy[i] = 0.0;
    for (int X = 0; X < nx; ++X) {
        for (int Y = (int)ceil(p->yr1[X] - 0.01); Y <= (int)floor(p->yr2[X] + 0.01); ++Y) {
            int i = XY(p, X, Y);
            double dpx = 0.0;
            if (X == 0) dpx = p->P[XY(p, 1, Y)] - p->P[i];
            else if (X == nx - 1) dpx = p->P[i] - p->P[XY(p, X - 1, Y)];
            else dpx = 0.5 * (p->P[XY(p, X + 1, Y)] - p->P[XY(p, X - 1, Y)]);
            double mu = p->Rho0 * (p->tau - 0.5) / 3.0;
            p->Ux[i] = dpx / (2.0 * mu) * ((Y - p->yr1[X]) * (Y - p->yr2[X]));
        }
    }
    for (int X = 0; X < nx; ++X)
        for (int Y = 0; Y < ny; ++Y) {
            int i = XY(p, X, Y);
            if (solid(p, X, Y)) { for (int k = 0; k < 9; ++k) GIN(p, i, k) = 0.0; continue; }
            double geq[9];
            equilibrium_g(p, p->P[i], p->Ux[i], p->Uy[i], geq);
            for (int k = 0; k < 9; ++k) GIN(p, i, k) = geq[k];
        }
}

/* AB:216-230 */
static void macroscopic(pulsatile *p)
{
    for (int X = 0; X < p->nx; ++X)
        for (int Y = 0; Y < p->ny; ++Y) {
            int i = XY(p, X, Y);
            if (solid(p, X, Y)) { p->P[i] = p->Ux[i] = p->Uy[i] = 0.0; continue; }
            double pp = 0.0, ux = 0.0, uy = 0.0;
            for (int k = 0; k < 9; ++k) pp += GIN(p, i, k);
            for (int k = 1; k < 9; ++k) { ux += GIN(p, i, k) * CK[k][0]; uy += GIN(p, i, k) * CK[k][1]; }
            p->P[i] = pp;
            p->Ux[i] = 3.0 * ux / p->Rho0;
            p->Uy[i] = 3.0 * uy / p->Rho0;
        }
}

/* AB:243-272 */
static void move_walls(pulsatile *p)
{
    const int nx = p->nx;
    const double Yw1 = 0.0, Yw2 = (double)(p->ny - 1);
    for (int Xf = 0; Xf < nx; ++Xf) {
        double Ps = p->P[XY(p, Xf, Y0c(p))] - p->p_tissue;
        double target = (Yw1 + 0.5) - Ps / p->alpha;
        double d = target - p->yr1[Xf];
        double cap = 0.25;
        if (d > cap) d = cap;
        if (d < -cap) d = -cap;
        p->y1new[Xf] = p->yr1[Xf] + d;
    }
    for (int X = 0; X < nx; ++X) { p->Vw1[X] = p->y1new[X] - p->yr1[X]; p->yr1[X] = p->y1new[X]; }
    for (int Xf = 0; Xf < nx; ++Xf) {
        double Ps = p->P[XY(p, Xf, Y0c(p) + 1)] - p->p_tissue;
        double target = (Yw2 - 0.5) + Ps / p->alpha;
        double d = target - p->yr2[Xf];
        double cap = 0.25;
        if (d > cap) d = cap;
        if (d < -cap) d = -cap;
        p->y2new[Xf] = p->yr2[Xf] + d;
    }
    for (int X = 0; X < nx; ++X) { p->Vw2[X] = p->y2new[X] - p->yr2[X]; p->yr2[X] = p->y2new[X]; }
}

/* AB:418-458 */
static void seed_from_nearest_fluid(pulsatile *p, int X, int Y)
{
    static const int dx[8] = {1, -1, 0, 0, 1, 1, -1, -1};
    static const int dy[8] = {0, 0, 1, -1, 1, -1, 1, -1};
    int i_dst = XY(p, X, Y), any = 0, cnt = 0;
    double acc[9] = {0};
    for (int n = 0; n < 8; ++n) {
        int Xn = X + dx[n], Yn = Y + dy[n];
        if (Xn < 0 || Xn >= p->nx || Yn < 0 || Yn >= p->ny) continue;
        if (solid(p, Xn, Yn)) continue;
        int i_src = XY(p, Xn, Yn);
        for (int k = 0; k < 9; ++k) acc[k] += GIN(p, i_src, k);
        any = 1; ++cnt;
    }
    for (int R = 2; !any && R <= 4; ++R)
        for (int sx = -R; sx <= R; ++sx) {
            int sy_top = R - abs(sx), sy_bot = -sy_top;
            for (int w = 0; w < 2; ++w) {
                int sy = w ? sy_bot : sy_top;
                int Xn = X + sx, Yn = Y + sy;
                if (Xn < 0 || Xn >= p->nx || Yn < 0 || Yn >= p->ny) continue;
                if (solid(p, Xn, Yn)) continue;
                int i_src = XY(p, Xn, Yn);
                for (int k = 0; k < 9; ++k) acc[k] += GIN(p, i_src, k);
                any = 1; ++cnt;
            }
        }
    if (any && cnt > 0) {
        for (int k = 0; k < 9; ++k) GIN(p, i_dst, k) = acc[k] / (double)cnt;
    } else {
        double geqv[9];
        equilibrium_g(p, p->P[i_dst], 0.0, 0.0, geqv);
        for (int k = 0; k < 9; ++k) GIN(p, i_dst, k) = geqv[k];
    }
}

/* AB:401-416, :460-498 */
static void fill_fluid_node(pulsatile *p, int X, int Y, int Ffrac[3][3])
{
    if (X == 0 || X == p->nx - 1) {
        int Ys = (Y < Y0c(p)) ? Y + 1 : Y - 1;
        int is = XY(p, X, Ys), id = XY(p, X, Y);
        for (int I = 0; I < 9; ++I) GIN(p, id, KFROMI[I]) = GIN(p, is, KFROMI[I]);
    } else {
        int SumFrac = 0;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) SumFrac += Ffrac[i][j];
        int id = XY(p, X, Y);
        if (SumFrac == 0) seed_from_nearest_fluid(p, X, Y);
        else
            for (int I = 0; I < 9; ++I)
                if (Ffrac[1 - EXI[I]][1 - EYI[I]] != 1) {
                    const int k = KFROMI[I];
                    double acc = 0.0;
                    acc += GIN(p, XY(p, X - 1, Y - 1), k) * Ffrac[0][0];
                    acc += GIN(p, XY(p, X, Y - 1), k) * Ffrac[1][0];
                    acc += GIN(p, XY(p, X + 1, Y - 1), k) * Ffrac[2][0];
                    acc += GIN(p, XY(p, X - 1, Y), k) * Ffrac[0][1];
                    acc += GIN(p, XY(p, X + 1, Y), k) * Ffrac[2][1];
                    acc += GIN(p, XY(p, X - 1, Y + 1), k) * Ffrac[0][2];
                    acc += GIN(p, XY(p, X, Y + 1), k) * Ffrac[1][2];
                    acc += GIN(p, XY(p, X + 1, Y + 1), k) * Ffrac[2][2];
                    GIN(p, id, k) = acc / (double)SumFrac;
                }
    }
    /* Fresh_Macroscopic_Values AB:489-498 */
    int i = XY(p, X, Y);
    double pp = 0, ux = 0, uy = 0;
    for (int I = 0; I < 9; ++I) pp += GIN(p, i, KFROMI[I]);
    for (int I = 1; I < 9; ++I) { ux += GIN(p, i, KFROMI[I]) * EXI[I]; uy += GIN(p, i, KFROMI[I]) * EYI[I]; }
    p->P[i] = pp;
    p->Ux[i] = 3 * ux / p->Rho0;
    p->Uy[i] = 3 * uy / p->Rho0;
}

/* AB:384-399 */
static void update_fobj(pulsatile *p)
{
    const int NX = p->nx, NY = p->ny;
    memcpy(p->Fold, p->Fobj, (size_t)(NX + 2) * (NY + 2) * sizeof(double));
    init_fobj(p);
    int c1 = 0, c2 = 0;
    for (int X = 1; X <= NX; ++X)
        for (int Y = 1; Y <= NY; ++Y) {
            if (p->Fold[X * (NY + 2) + Y] < 1 && FO(p, X, Y) >= 1) {
                ++c1;
                int Ffrac[3][3];
                for (int i = -1; i <= 1; ++i)
                    for (int j = -1; j <= 1; ++j) Ffrac[i + 1][j + 1] = (int)(p->Fold[(X + i) * (NY + 2) + (Y + j)]);
                fill_fluid_node(p, X - 1, Y - 1, Ffrac);
            }
            if (p->Fold[X * (NY + 2) + Y] >= 1 && FO(p, X, Y) < 1) ++c2;
        }
    p->FreshNodes = c1;
    p->KilledNodes = c2;
}

/* AB:553-601 */
static void bouzidi(pulsatile *p, const border_node *B, int nb)
{
    const int nx = p->nx, ny = p->ny;
#define INDOM(Xp, Yp) ((Xp) >= 0 && (Xp) < nx && (Yp) >= 0 && (Yp) < ny)
    for (int b_ = 0; b_ < nb; ++b_) {
        int X = B[b_].X, Y = B[b_].Y;
        if (!INDOM(X, Y)) continue;
        for (int I = 1; I <= 8; ++I) {
            double D = B[b_].Delta[I - 1];
            if (D >= 1.0) continue;
            int jI = JBI[I], kI = KFROMI[I], kJ = KFROMI[jI];
            int X1 = X + EXI[I], Y1 = Y + EYI[I];
            int X2 = X1 + EXI[I], Y2 = Y1 + EYI[I];
            int X3 = X2 + EXI[I], Y3 = Y2 + EYI[I];
            if (!INDOM(X1, Y1)) continue;
            if (!INDOM(X2, Y2)) { X2 = X1; Y2 = Y1; }
            if (!INDOM(X3, Y3)) { X3 = X1; Y3 = Y1; }
            if (!INDOM(X3, Y3)) { X3 = X2; Y3 = Y2; }
            if (FO(p, X2 + 1, Y2 + 1) < 1) { X2 = X1; Y2 = Y1; }
            if (FO(p, X3 + 1, Y3 + 1) < 1) { X3 = X2; Y3 = Y2; }
            int b = XY(p, X, Y), n1 = XY(p, X1, Y1), n2 = XY(p, X2, Y2), n3 = XY(p, X3, Y3);
            if (D < 0.5) {
                GOUT(p, b, kI) = GOUT(p, n1, kJ) * (1 + 2 * D) * D + GOUT(p, n2, kJ) * (1 - 2 * D) * (1 + 2 * D) -
                                 GOUT(p, n3, kJ) * (1 - 2 * D) * D;
            } else {
                GOUT(p, b, kI) = (GOUT(p, n1, kJ) - GOUT(p, n1, kI) * (1 - 2 * D) * (1 + 2 * D) + GOUT(p, n2, kI) * (1 - 2 * D) * D) /
                                 (D * (1 + 2 * D));
            }
        }
    }
#undef INDOM
}

/* AB:603-616; wrapX periodic, wrapY identity (flat index spills into the neighbouring column, B.3) */
static void streaming(pulsatile *p)
{
    double tmp[9];
    for (int X = 0; X < p->nx; ++X)
        for (int Y = 0; Y < p->ny; ++Y) {
            int i = XY(p, X, Y);
            for (int k = 0; k < 9; ++k) {
                int XX = (X - CK[k][0] + p->nx) % p->nx;
                int YY = Y - CK[k][1];
                long src = (long)YY + (long)p->ny * XX;
                tmp[k] = GOUT(p, src, k);
            }
            for (int k = 0; k < 9; ++k) GIN(p, i, k) = tmp[k];
        }
}

/* AB:618-669 */
static void zou_he(pulsatile *p, int t_iter)
{
    const double Rho0 = p->Rho0;
    double Pin = p->p0_in;
    if (t_iter >= p->t_start) Pin = p->p0_in + p->p_oscillatory * sin(p->omega * (t_iter + 1 - p->t_start));
    {
        int X = 0, ylo = (int)ceil(p->yr1[0] - 0.01), yhi = (int)floor(p->yr2[0] + 0.01);
        if (ylo < 0) ylo = 0;
        if (yhi > p->ny - 1) yhi = p->ny - 1;
        for (int Y = ylo; Y <= yhi; ++Y) {
            int i = XY(p, X, Y);
            double g0 = GIN(p, i, KFROMI[0]), g2 = GIN(p, i, KFROMI[2]), g3 = GIN(p, i, KFROMI[3]);
            double g4 = GIN(p, i, KFROMI[4]), g6 = GIN(p, i, KFROMI[6]), g7 = GIN(p, i, KFROMI[7]);
            double Uin = Pin - g0 - g2 - 2 * g3 - g4 - 2 * g6 - 2 * g7;
            Uin = Uin * 3.0 / Rho0;
            GIN(p, i, KFROMI[1]) = g3 + 2.0 * Rho0 / 9.0 * Uin;
            GIN(p, i, KFROMI[5]) = Rho0 / 18.0 * Uin - 0.5 * (g2 - g4) + g7;
            GIN(p, i, KFROMI[8]) = Rho0 / 18.0 * Uin + 0.5 * (g2 - g4) + g6;
        }
    }
    double Pout = p->p0_out;
    if (t_iter >= p->t_start + p->t_propagation)
        Pout = p->p0_out + p->p_oscillatory * sin(p->omega * (t_iter + 1 - p->t_start - p->t_propagation));
    if (t_iter > p->t_sever) Pout = 0;
    {
        int X = p->nx - 1, ylo = (int)ceil(p->yr1[p->nx - 1] - 0.01), yhi = (int)floor(p->yr2[p->nx - 1] + 0.01);
        if (ylo < 0) ylo = 0;
        if (yhi > p->ny - 1) yhi = p->ny - 1;
        for (int Y = ylo; Y <= yhi; ++Y) {
            int i = XY(p, X, Y);
            double g0 = GIN(p, i, KFROMI[0]), g1 = GIN(p, i, KFROMI[1]), g2 = GIN(p, i, KFROMI[2]);
            double g4 = GIN(p, i, KFROMI[4]), g5 = GIN(p, i, KFROMI[5]), g8 = GIN(p, i, KFROMI[8]);
            double Uout = g0 + 2 * g1 + g2 + g4 + 2 * g5 + 2 * g8 - Pout;
            Uout = Uout * 3.0 / Rho0;
            GIN(p, i, KFROMI[3]) = g1 - 2.0 * Rho0 / 9.0 * Uout;
            GIN(p, i, KFROMI[6]) = -Rho0 / 18.0 * Uout - 0.5 * (g2 - g4) + g8;
            GIN(p, i, KFROMI[7]) = -Rho0 / 18.0 * Uout + 0.5 * (g2 - g4) + g5;
        }
    }
}

/* ---- public API (ctypes) ------------------------------------------------------------------ */
void pulsatile_destroy(pulsatile *p)
{
    if (!p) return;
    free(p->lattice); free(p->flag); free(p->P); free(p->Ux); free(p->Uy); free(p->yr1); free(p->yr2);
    free(p->y1new); free(p->y2new); free(p->Vw1); free(p->Vw2); free(p->Fobj); free(p->Fold); free(p->B1); free(p->B2);
    free(p);
}

/* driver set-up AB:719-757 with N a parameter (the reference hard-codes N = 64); nx = 1 + 10 (N-2), ny = N */
pulsatile *pulsatile_create(int N, double tau, double alpha, double p0_in, double p0_out, int is_severed, int deformable)
{
    pulsatile *p = (pulsatile *)calloc(1, sizeof(pulsatile));
    p->nx = 1 + 10 * (N - 2);
    p->ny = N;
    p->nelem = (size_t)p->nx * p->ny;
    p->npop = 9 * p->nelem;
    p->lattice = (double *)calloc(2 * p->npop + 1, sizeof(double));
    p->flag = (uint8_t *)malloc(p->nelem);
    memset(p->flag, 1, p->nelem);
    p->P = (double *)calloc(p->nelem, sizeof(double));
    p->Ux = (double *)calloc(p->nelem, sizeof(double));
    p->Uy = (double *)calloc(p->nelem, sizeof(double));
    p->yr1 = (double *)calloc(p->nx, sizeof(double)); p->yr2 = (double *)calloc(p->nx, sizeof(double));
    p->y1new = (double *)calloc(p->nx, sizeof(double)); p->y2new = (double *)calloc(p->nx, sizeof(double));
    p->Vw1 = (double *)calloc(p->nx, sizeof(double)); p->Vw2 = (double *)calloc(p->nx, sizeof(double));
    const size_t nf = (size_t)(p->nx + 2) * (p->ny + 2);
    p->Fobj = (double *)malloc(nf * sizeof(double));
    p->Fold = (double *)malloc(nf * sizeof(double));
    for (size_t i = 0; i < nf; ++i) p->Fobj[i] = 1.0;
    p->Rho0 = 1.0 / pow(1, 3);
    p->tau = tau; p->s8 = 1.0 / tau; p->s5 = 1.0;
    const double S[9] = {1, 1, 1, 1, p->s5, 1, p->s5, p->s8, p->s8};
    memcpy(p->S, S, sizeof(S));
    p->deformable = deformable; p->is_severed = is_severed; p->alpha = alpha;
    p->p0_in = p0_in; p->p0_out = p0_out;
    p->t_beat = p->nx > 1 ? p->nx : 1;
    setup_parameters(p);
    if (theoretical_wall_and_pressure(p)) { pulsatile_destroy(p); return NULL; }
    memcpy(p->y1new, p->yr1, p->nx * sizeof(double));
    memcpy(p->y2new, p->yr2, p->nx * sizeof(double));
    init_fobj(p);
    update_boundary_bottom(p);
    update_boundary_top(p);
    initialize_P_U_g(p);
    (void)X0c;
    return p;
}

/* n iterations of the reference loop body AB:764-790 (without the VTK dump) */
void pulsatile_step(pulsatile *p, int n)
{
    for (int s = 0; s < n; ++s) {
        const int t = p->t_iter;
        for (int X = 0; X < p->nx; ++X) for (int Y = 0; Y < p->ny; ++Y) mrt_collision(p, X, Y);
        bouzidi(p, p->B1, p->Nb1);
        bouzidi(p, p->B2, p->Nb2);
        streaming(p);
        zou_he(p, t);
        macroscopic(p);
        if (p->deformable) {
            move_walls(p);
            update_fobj(p);
            update_boundary_bottom(p);
            update_boundary_top(p);
        }
        p->parity = 1 - p->parity;
        p->t_iter++;
    }
}

/* replace the whole functor state (host arrays in the reference layout) and rebuild what the reference derives from the
 * wall positions: Fobj + mask (AB:275-285) and the border lists (AB:294-382).  Used to start from states other than
 * Initialize_P_U_g, e.g. the open vessel at rest of bench.py / tests. */
int pulsatile_set_state(pulsatile *p, const double *lattice, const double *P, const double *Ux, const double *Uy,
                        const double *yr1, const double *yr2, int parity, int t_iter)
{
    memcpy(p->lattice, lattice, 2 * p->npop * sizeof(double));
    memcpy(p->P, P, p->nelem * sizeof(double));
    memcpy(p->Ux, Ux, p->nelem * sizeof(double));
    memcpy(p->Uy, Uy, p->nelem * sizeof(double));
    memcpy(p->yr1, yr1, p->nx * sizeof(double));
    memcpy(p->yr2, yr2, p->nx * sizeof(double));
    memcpy(p->y1new, yr1, p->nx * sizeof(double));
    memcpy(p->y2new, yr2, p->nx * sizeof(double));
    p->parity = parity;
    p->t_iter = t_iter;
    init_fobj(p);
    update_boundary_bottom(p);
    update_boundary_top(p);
    return 0;
}

int pulsatile_nx(const pulsatile *p) { return p->nx; }
int pulsatile_ny(const pulsatile *p) { return p->ny; }
int pulsatile_parity(const pulsatile *p) { return p->parity; }
int pulsatile_tf(const pulsatile *p) { return p->t_beat + 2 * p->t_propagation; }
int pulsatile_fresh(const pulsatile *p) { return p->FreshNodes; }
const double *pulsatile_lattice(const pulsatile *p) { return p->lattice; }
void pulsatile_get(const pulsatile *p, double *P, double *Ux, double *Uy, uint8_t *flag, double *yr1, double *yr2)
{
    if (P) memcpy(P, p->P, p->nelem * sizeof(double));
    if (Ux) memcpy(Ux, p->Ux, p->nelem * sizeof(double));
    if (Uy) memcpy(Uy, p->Uy, p->nelem * sizeof(double));
    if (flag) memcpy(flag, p->flag, p->nelem);
    if (yr1) memcpy(yr1, p->yr1, p->nx * sizeof(double));
    if (yr2) memcpy(yr2, p->yr2, p->nx * sizeof(double));
}

/* legacy-VTK text of P/Ux/Uy/Flag arrays exactly as saveVtkFields_PulsatileBloodFlow2D prints them (AB:680-706):
 * `os << float(x)` with the default stream format == printf("%g") of the float value */
int pulsatile_write_vtk(int nx, int ny, const double *P, const double *Ux, const double *Uy, const uint8_t *flag,
                        int time_iter, const char *path)
{
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    const double dx = 1.0 / nx;
    fprintf(f, "# vtk DataFile Version 2.0\niteration %d\nASCII\n\nDATASET STRUCTURED_POINTS\n", time_iter);
    fprintf(f, "DIMENSIONS %d %d %d\nORIGIN 0 0 0\nSPACING %g %g %g\n\nPOINT_DATA %d\n", nx, ny, 1, dx, dx, dx, nx * ny);
    const double *arr[3] = {P, Ux, Uy};
    const char *nm[3] = {"P", "Ux", "Uy"};
    for (int a = 0; a < 3; ++a) {
        fprintf(f, "SCALARS %s float 1\nLOOKUP_TABLE default\n", nm[a]);
        for (int y = 0; y < ny; ++y) {
            for (int x = 0; x < nx; ++x) fprintf(f, "%g ", (double)(float)arr[a][y + ny * x]);
            fputc('\n', f);
        }
        fputc('\n', f);
    }
    fprintf(f, "SCALARS Flag int 1\nLOOKUP_TABLE default\n");
    for (int y = 0; y < ny; ++y) {
        for (int x = 0; x < nx; ++x) fprintf(f, "%d ", flag[y + ny * x] == 0 ? 1 : 0);
        fputc('\n', f);
    }
    fputc('\n', f);
    fclose(f);
    return 0;
}
