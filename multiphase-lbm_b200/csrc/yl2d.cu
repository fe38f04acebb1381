// yl2d.cu -- "Abbashub LBM/apps/Young_Laplace2D.h" (AB/ below) on the device: conservative phase-field LBM (Fakhari et
// al. 2017: population h for the phase field, velocity-based population g for pressure / momentum), D2Q9 BGK, fully
// periodic.  It is the problem the AB reference build runs by default (AB/apps/COOLBM.cpp:99).
//
// The reference splits an iteration into collide_stream_at (parallel, AB:217-290, reads ten STORED fields) and
// update_fields (serial, five sweeps over the lattice, AB:297-370).  Here the stored fields disappear: update_fields of
// iteration t is evaluated at the start of iteration t+1 inside the collide kernel, from the populations and the 3x3
// neighbourhood of phi, so that one iteration is
//   yl2d_phi     phi = sum_k h_k                                        9 reads + 1 write per node
//   yl2d_step    rho, grad phi, lap phi -> mu, n, p*, u (AB:297-370) then collide + push of h and g (AB:217-290)
//                18 populations + u_prev + phi(3x3, cached) read, 18 populations + u written
// (416 B per node against ~700 B for the reference's field-by-field form).  The only state besides the populations is
// the velocity of the previous update (the viscous force of AB:353-358 is evaluated with it).
// Device layout == reference layout: lattice[h_in | h_out | g_in | g_out] selected by parity, i = y + ny x.
//
// Built with -fmad=false and written in the reference's operation order: results are BIT-IDENTICAL to the reference.
// That is a necessity, not a nicety: the model amplifies rounding-level differences (the unit normal grad phi/|grad phi|
// far from the interface) by ~100x per 50 iterations until they saturate near 1e-5, so no FMA-contracted build can
// hold 1e-10 for 1000 iterations (measured).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "clbm_internal.h"

namespace clbm {
namespace yl {

struct Geo { int nx, ny; long long ne; };
struct Par { double Rhol, Rhoh, Sigma, W, M, tau, s8, Beta, kappa, dRho3, wc, fac, rW; };   // fac = (0.5 - tau)/tau, rW = RN(1/W)

// Correctly rounded quotients without the ~30-instruction IEEE division sequence (Markstein): with r = RN(1/d),
// q0 = RN(x r), rem = x - d q0 (exact, FMA), q = RN(q0 + rem r) equals RN(x / d) for normal quotients.  divc: compile-time
// divisor; divr: run-time divisor whose correctly rounded reciprocal is shared by several quotients (one true division
// per node instead of twenty).  Checked against x / d on 3e8 / 4e8 random operands each: no mismatch.
template <int C> __device__ __forceinline__ double divc(double x)
{
    constexpr double r = 1.0 / C;
    const double q0 = __dmul_rn(x, r);
    return __fma_rn(__fma_rn(-(double)C, q0, x), r, q0);
}
__device__ __forceinline__ double divr(double x, double d, double r)
{
    const double q0 = __dmul_rn(x, r);
    return __fma_rn(__fma_rn(-d, q0, x), r, q0);
}

__host__ __device__ constexpr int ckx(int k) { constexpr int v[9] = {-1, 0, -1, -1, 0, 1, 0, 1, 1}; return v[k]; }
__host__ __device__ constexpr int cky(int k) { constexpr int v[9] = {0, -1, -1, 1, 0, 0, 1, 1, -1}; return v[k]; }
__host__ __device__ constexpr double tk(int k)
{
    constexpr double v[9] = {1. / 9., 1. / 9., 1. / 36., 1. / 36., 4. / 9., 1. / 9., 1. / 9., 1. / 36., 1. / 36.};
    return v[k];
}

// AB:174-180
__device__ __forceinline__ void gawa(double U, double V, double out[9])
{
    const double U2 = U * U + V * V;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const double eU = ckx(k) * U + cky(k) * V;
        out[k] = tk(k) * (3.0 * eU + 4.5 * eU * eU - 1.5 * U2);
    }
}
// AB:183-201
__device__ __forceinline__ void viscous_force(const Par &p, double dcdx, double dcdy, const double gneq[9], double &FmX, double &FmY)
{
    double sxx = 0.0, sxy = 0.0, syy = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (k == 4) continue;
        sxx += gneq[k] * (ckx(k) * ckx(k));
        sxy += gneq[k] * (ckx(k) * cky(k));
        syy += gneq[k] * (cky(k) * cky(k));
    }
    const double fac = p.fac, dR = p.Rhoh - p.Rhol;
    FmX = fac * (sxx * dcdx + sxy * dcdy) * dR;
    FmY = fac * (sxy * dcdx + syy * dcdy) * dR;
}

// what update_fields (AB:297-370) leaves at one node
struct Node { double C, Rho, P, Ux, Uy, dCx, dCy, mu, ni, nj, den, rden; };   // den = Rho + 1e-30, rden = RN(1 / den)

// update_fields (AB:297-370) at one node from the 3x3 neighbourhood of phi, the g populations of the node and the previous velocity
__device__ __forceinline__ void node_from_stencil(const Par &p, double cC, double cE, double cW, double cN, double cS, double cNE,
                                                  double cNW, double cSE, double cSW, const double gin[9], double Uo, double Vo,
                                                  bool keep_u, Node &n)
{
    n.C = cC;
    n.Rho = p.Rhol + cC * (p.Rhoh - p.Rhol);
    n.dCx = divc<3>(cE - cW) + divc<12>(cSE + cNE - cSW - cNW);
    n.dCy = divc<3>(cN - cS) + divc<12>(cNW + cNE - cSW - cSE);
    const double D2C = divc<6>(cSW + cSE + cNW + cNE + 4.0 * (cS + cW + cE + cN) - 20.0 * cC);
    n.den = n.Rho + 1e-30;
    n.rden = 1.0 / n.den;
    n.mu = 4.0 * p.Beta * cC * (cC - 1.0) * (cC - 0.5) - p.kappa * D2C;
    const double inv = 1.0 / sqrt(n.dCx * n.dCx + n.dCy * n.dCy + 1e-32);
    n.ni = n.dCx * inv;
    n.nj = n.dCy * inv;
    double pstar = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) pstar += gin[k];
    n.P = pstar;
    if (keep_u) { n.Ux = Uo; n.Uy = Vo; return; }      // the uploaded velocity is already the updated one
    const double FpX = -n.P * p.dRho3 * n.dCx, FpY = -n.P * p.dRho3 * n.dCy;
    double GaWa[9], gneq[9];
    gawa(Uo, Vo, GaWa);
#pragma unroll
    for (int k = 0; k < 9; ++k) gneq[k] = gin[k] - (n.P * tk(k) + GaWa[k]);
    double FmX, FmY;
    viscous_force(p, n.dCx, n.dCy, gneq, FmX, FmY);
    const double Fx = n.mu * n.dCx + FpX + FmX, Fy = n.mu * n.dCy + FpY + FmY;
    double mx = 0.0, my = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) { mx += gin[k] * ckx(k); my += gin[k] * cky(k); }
    n.Ux = mx + divr(0.5 * Fx, n.den, n.rden);
    n.Uy = my + divr(0.5 * Fy, n.den, n.rden);
}

__device__ __forceinline__ void update_node(const Par &p, const Geo &g, const double *__restrict__ C, int X, int Y, const double gin[9],
                                            double Uo, double Vo, bool keep_u, Node &n)
{
    const int xm = X == 0 ? g.nx - 1 : X - 1, xp = X == g.nx - 1 ? 0 : X + 1;
    const int ym = Y == 0 ? g.ny - 1 : Y - 1, yp = Y == g.ny - 1 ? 0 : Y + 1;
    auto at = [&](int x, int y) { return C[y + (long long)g.ny * x]; };
    node_from_stencil(p, at(X, Y), at(xp, Y), at(xm, Y), at(X, yp), at(X, ym), at(xp, yp), at(xm, yp), at(xp, ym), at(xm, ym), gin, Uo, Vo,
                      keep_u, n);
}

// collide_stream_at (AB:217-290) of one node whose fields are in n
__device__ __forceinline__ void collide_push(const Par &p, const Geo &g, int X, int Y, long long i, const Node &n, const double hk[9],
                                             const double gin[9], double *__restrict__ hout, double *__restrict__ gout)
{
    double GaWa[9];
    gawa(n.Ux, n.Uy, GaWa);
    const double shape = divr(1.0 - 4.0 * (n.C - 0.5) * (n.C - 0.5), p.W, p.rW);
    const double FpX = -n.P * p.dRho3 * n.dCx, FpY = -n.P * p.dRho3 * n.dCy;
    double gneq[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) gneq[k] = gin[k] - (n.P * tk(k) + GaWa[k]);
    double FmX, FmY;
    viscous_force(p, n.dCx, n.dCy, gneq, FmX, FmY);
    const double Fx = n.mu * n.dCx + FpX + FmX, Fy = n.mu * n.dCy + FpY + FmY;
    const int xm = X == 0 ? g.nx - 1 : X - 1, xp = X == g.nx - 1 ? 0 : X + 1;
    const int ym = Y == 0 ? g.ny - 1 : Y - 1, yp = Y == g.ny - 1 ? 0 : Y + 1;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const double hlp_h = tk(k) * (shape * (ckx(k) * n.ni + cky(k) * n.nj));
        const double heq = n.C * (tk(k) + GaWa[k]) - 0.5 * hlp_h;
        const double hlp_g = divr(3.0 * tk(k) * (ckx(k) * Fx + cky(k) * Fy), n.den, n.rden);
        const double geq = (n.P * tk(k) + GaWa[k]) - 0.5 * hlp_g;
        const double ho = (1.0 - p.wc) * hk[k] + p.wc * heq + hlp_h;
        const double go = (1.0 - p.s8) * gin[k] + p.s8 * geq + hlp_g;
        const long long nb = (ckx(k) < 0 ? xm : (ckx(k) > 0 ? xp : X)) * (long long)g.ny + (cky(k) < 0 ? ym : (cky(k) > 0 ? yp : Y));
        hout[k * g.ne + nb] = ho;
        gout[k * g.ne + nb] = go;
    }
}

__global__ void __launch_bounds__(256) yl2d_phi(const double *__restrict__ hin, double *__restrict__ C, Geo g)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.ne) return;
    double phi = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) phi += hin[k * g.ne + i];
    C[i] = phi;
}

// update_fields of the previous iteration + collide_stream_at of this one (AB:217-290)
__global__ void __launch_bounds__(256) yl2d_step(const double *__restrict__ hin, double *__restrict__ hout, const double *__restrict__ gin_,
                                                 double *__restrict__ gout, const double *__restrict__ C, double *__restrict__ Ux,
                                                 double *__restrict__ Uy, Geo g, Par p, int keep_u)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.ne) return;
    const int X = (int)(i / g.ny), Y = (int)(i % g.ny);
    double gin[9], hk[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { gin[k] = gin_[k * g.ne + i]; hk[k] = hin[k * g.ne + i]; }
    Node n;
    update_node(p, g, C, X, Y, gin, Ux[i], Uy[i], keep_u != 0, n);
    Ux[i] = n.Ux;
    Uy[i] = n.Uy;
    collide_push(p, g, X, Y, i, n, hk, gin, hout, gout);
}

// Fused form of yl2d_phi + yl2d_step: a CTA owns NT-2 consecutive rows and marches along x with a 3-column shared-memory
// ring of phi; the h populations are loaded once (two columns ahead, in registers) and serve both phi and the collision
// one column later, so the separate phi pass (9 reads + 1 write per node) disappears.  Arithmetic is yl2d_step's, operation
// for operation (same update_node / collision code), hence still bit-identical to the reference.
template <int NT>
__global__ void __launch_bounds__(NT) yl2d_fused(const double *__restrict__ hin, double *__restrict__ hout, const double *__restrict__ gin_,
                                                 double *__restrict__ gout, double *__restrict__ Ux, double *__restrict__ Uy, Geo g, Par p,
                                                 int keep_u, int xchunk)
{
    __shared__ double Cr[3][NT];
    const int tid = threadIdx.x;
    const int Yraw = (int)blockIdx.x * (NT - 2) - 1 + tid;
    const bool row_ok = Yraw >= -1 && Yraw <= g.ny;                      // rows -1 and ny are the periodic images
    const int Y = Yraw < 0 ? Yraw + g.ny : (Yraw >= g.ny ? Yraw - g.ny : Yraw);
    const bool own = tid >= 1 && tid < NT - 1 && Yraw < g.ny;
    const int xa = blockIdx.y * xchunk, xb = min(g.nx, xa + xchunk);
    auto wrapx = [&](int X) { return X < 0 ? X + g.nx : (X >= g.nx ? X - g.nx : X); };
    auto load_h = [&](int X, double h[9]) {
        if (!row_ok) return;
        const long long i = Y + (long long)g.ny * wrapx(X);
#pragma unroll
        for (int k = 0; k < 9; ++k) h[k] = hin[k * g.ne + i];
    };
    auto put_phi = [&](int X, const double h[9]) {
        double phi = 0.0;
#pragma unroll
        for (int k = 0; k < 9; ++k) phi += h[k];
        Cr[(X + 3) % 3][tid] = phi;
    };
    double hc[9], hn[9], hp[9];          // columns X, X+1, X+2 (prefetch)
    load_h(xa - 1, hp);
    put_phi(xa - 1, hp);
    load_h(xa, hc);
    put_phi(xa, hc);
    load_h(xa + 1, hn);
    for (int X = xa; X < xb; ++X) {
        if (X + 1 < xb) load_h(X + 2, hp);          // in flight while this column is processed
        put_phi(X + 1, hn);
        __syncthreads();
        if (own) {
            const long long i = Y + (long long)g.ny * X;
            double gin[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) gin[k] = gin_[k * g.ne + i];
            // update_fields at this node with phi from the ring (same expressions as update_node)
            const int sm = (X + 2) % 3, s0 = X % 3, sp = (X + 1) % 3;
            Node n;
            {
                const double cC = Cr[s0][tid], cE = Cr[sp][tid], cW = Cr[sm][tid], cN = Cr[s0][tid + 1], cS = Cr[s0][tid - 1];
                const double cNE = Cr[sp][tid + 1], cNW = Cr[sm][tid + 1], cSE = Cr[sp][tid - 1], cSW = Cr[sm][tid - 1];
                node_from_stencil(p, cC, cE, cW, cN, cS, cNE, cNW, cSE, cSW, gin, Ux[i], Uy[i], keep_u != 0, n);
            }
            Ux[i] = n.Ux;
            Uy[i] = n.Uy;
            collide_push(p, g, X, Y, i, n, hc, gin, hout, gout);
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) { hc[k] = hn[k]; hn[k] = hp[k]; }
        __syncthreads();
    }
}

// the fields update_fields would hold now (download / diagnostics); does not touch the stored velocity
__global__ void __launch_bounds__(256) yl2d_fields(const double *__restrict__ gin_, const double *__restrict__ C, const double *__restrict__ Ux,
                                                   const double *__restrict__ Uy, Geo g, Par p, int keep_u, double *oC, double *oP,
                                                   double *oRho, double *oUx, double *oUy)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.ne) return;
    double gin[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) gin[k] = gin_[k * g.ne + i];
    Node n;
    update_node(p, g, C, (int)(i / g.ny), (int)(i % g.ny), gin, Ux[i], Uy[i], keep_u != 0, n);
    if (oC) oC[i] = n.C;
    if (oP) oP[i] = n.P;
    if (oRho) oRho[i] = n.Rho;
    if (oUx) oUx[i] = n.Ux;
    if (oUy) oUy[i] = n.Uy;
}

// iniCell (AB:141-169) on the host, with the reference's expressions and libm (tanh, sqrt): the model amplifies
// rounding differences by ~100x per 50 iterations, so the initial state must be the reference's to the bit
static void init_host(const Geo &g, const Par &p, std::vector<double> &lat)
{
    const size_t ne = (size_t)g.ne;
    for (size_t i = 0; i < ne; ++i) {
        const int X = (int)(i / g.ny), Y = (int)(i % g.ny);
        const double xc = double(g.nx) / 2.0 - 0.5, yc = double(g.ny) / 2.0 - 0.5, R0 = double(g.nx) / 8.0;
        const double r = std::sqrt((X - xc) * (X - xc) + (Y - yc) * (Y - yc));
        const double phi = 0.5 - 0.5 * std::tanh(2.0 * (R0 - r) / p.W);
        const double rho = p.Rhol + phi * (p.Rhoh - p.Rhol);
        double P = 0.0;
        const double prho = (rho + 1e-12) / 3.0;
        const double corr = (phi * p.Sigma / R0) / prho;
        P -= corr;
        for (int k = 0; k < 9; ++k) { lat[(size_t)k * ne + i] = phi * tk(k); lat[18 * ne + (size_t)k * ne + i] = P * tk(k); }
    }
}

// deterministic two-stage sums: a = sum x, b = sum (y^2 + z^2)
__global__ void __launch_bounds__(256) yl2d_reduce1(const double *x, const double *y, const double *z, long long n, double *part)
{
    double a = 0.0, b = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        a += x[i];
        b += y[i] * y[i] + z[i] * z[i];
    }
    __shared__ double sa[8], sb[8];
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { a += sa[w]; b += sb[w]; }
        part[2 * blockIdx.x] = a;
        part[2 * blockIdx.x + 1] = b;
    }
}
__global__ void yl2d_reduce2(const double *part, int nb, double *out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int j = 0; j < nb; ++j) { a += part[2 * j]; b += part[2 * j + 1]; }
        out[0] = a;
        out[1] = b;
    }
}

}  // namespace yl
}  // namespace clbm

using namespace clbm;
using namespace clbm::yl;

struct clbm_yl2d {
    clbm_yl2d_params prm;
    Geo g;
    Par p;
    int device, parity, keep_u;    // keep_u: the stored velocity is already the updated one (right after an upload)
    long long steps;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    int64_t launches;
    double *lat, *C, *Ux, *Uy, *tmp, *part, *red_host;
};

namespace {
constexpr int RED_BLOCKS = 592;

double *h_in(clbm_yl2d *c) { return c->lat + (size_t)c->parity * 9 * c->g.ne; }
double *h_out(clbm_yl2d *c) { return c->lat + (size_t)(1 - c->parity) * 9 * c->g.ne; }
double *g_in(clbm_yl2d *c) { return c->lat + (size_t)18 * c->g.ne + (size_t)c->parity * 9 * c->g.ne; }
double *g_out(clbm_yl2d *c) { return c->lat + (size_t)18 * c->g.ne + (size_t)(1 - c->parity) * 9 * c->g.ne; }

int one_step(clbm_yl2d *c)
{
    static const int fused = getenv("CLBM_YL2D_FUSED") ? atoi(getenv("CLBM_YL2D_FUSED")) : 1;
    if (fused && c->g.ny >= 3) {
        constexpr int NT = 128;
        int xchunk = c->g.nx < 32 ? c->g.nx : 32;
        if (const char *e = getenv("CLBM_YL2D_XCHUNK")) { const int v = atoi(e); if (v > 0) xchunk = v < c->g.nx ? v : c->g.nx; }
        dim3 grid((c->g.ny + (NT - 2) - 1) / (NT - 2), (c->g.nx + xchunk - 1) / xchunk);
        yl2d_fused<NT><<<grid, NT, 0, c->stream>>>(h_in(c), h_out(c), g_in(c), g_out(c), c->Ux, c->Uy, c->g, c->p, c->keep_u, xchunk);
        c->launches += 1;
    } else {
        const int nb = grid_for(c->g.ne, 256);
        yl2d_phi<<<nb, 256, 0, c->stream>>>(h_in(c), c->C, c->g);
        yl2d_step<<<nb, 256, 0, c->stream>>>(h_in(c), h_out(c), g_in(c), g_out(c), c->C, c->Ux, c->Uy, c->g, c->p, c->keep_u);
        c->launches += 2;
    }
    c->keep_u = 0;
    c->parity = 1 - c->parity;
    c->steps++;
    CLBM_CUDA(cudaGetLastError());
    return CLBM_OK;
}

// current fields into c->tmp[0..4] = C, P, Rho, Ux, Uy
int current_fields(clbm_yl2d *c)
{
    const int nb = grid_for(c->g.ne, 256);
    const size_t ne = (size_t)c->g.ne;
    yl2d_phi<<<nb, 256, 0, c->stream>>>(h_in(c), c->C, c->g);
    yl2d_fields<<<nb, 256, 0, c->stream>>>(g_in(c), c->C, c->Ux, c->Uy, c->g, c->p, c->keep_u, c->tmp, c->tmp + ne, c->tmp + 2 * ne,
                                           c->tmp + 3 * ne, c->tmp + 4 * ne);
    c->launches += 2;
    CLBM_CUDA(cudaGetLastError());
    return CLBM_OK;
}
}  // namespace

extern "C" {

int clbm_yl2d_create(const clbm_yl2d_params *p, clbm_yl2d **out)
{
    if (!p || !out) { set_error("null argument"); return CLBM_EINVAL; }
    *out = nullptr;
    if (p->abi_version != CLBM_ABI_VERSION) { set_error("ABI version %d != %d", p->abi_version, CLBM_ABI_VERSION); return CLBM_EINVAL; }
    if (p->nx < 3 || p->ny < 3) { set_error("bad extent %d x %d", p->nx, p->ny); return CLBM_EINVAL; }
    if (!(p->tau > 0.0) || !(p->W > 0.0)) { set_error("bad tau / W"); return CLBM_EINVAL; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: this library has no CPU fallback"); return CLBM_ENODEVICE; }
    clbm_yl2d *c = new (std::nothrow) clbm_yl2d();
    if (!c) { set_error("out of host memory"); return CLBM_ENOMEM; }
    c->prm = *p;
    c->g.nx = p->nx; c->g.ny = p->ny; c->g.ne = (long long)p->nx * p->ny;
    c->device = p->device;
    if (c->device < 0) cudaGetDevice(&c->device);
    if (c->device >= ndev) { set_error("device %d of %d", c->device, ndev); delete c; return CLBM_EINVAL; }
    Par &q = c->p;     // derived constants as the driver sets them (AB:512-517)
    q.Sigma = p->Sigma; q.W = p->W; q.M = p->M; q.Rhol = p->RhoL; q.Rhoh = p->RhoH; q.tau = p->tau; q.s8 = 1.0 / p->tau;
    q.Beta = 12.0 * q.Sigma / q.W;
    q.kappa = 1.5 * q.Sigma * q.W;
    q.dRho3 = (q.Rhoh - q.Rhol) / 3.0;
    q.wc = 1.0 / (0.5 + 3.0 * q.M);
    q.fac = (0.5 - q.tau) / q.tau;
    q.rW = 1.0 / q.W;
    auto fail = [&](cudaError_t e, const char *what) { int r = cuda_fail(e, what, __FILE__, __LINE__); clbm_yl2d_destroy(c); return r; };
    cudaError_t e;
    const size_t ne = (size_t)c->g.ne;
    if ((e = cudaSetDevice(c->device)) != cudaSuccess) return fail(e, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    if ((e = cudaMalloc(&c->lat, 36 * ne * sizeof(double))) != cudaSuccess) return fail(e, "cudaMalloc lattice");
    if ((e = cudaMalloc(&c->C, ne * sizeof(double))) != cudaSuccess) return fail(e, "cudaMalloc C");
    if ((e = cudaMalloc(&c->Ux, ne * sizeof(double))) != cudaSuccess) return fail(e, "cudaMalloc Ux");
    if ((e = cudaMalloc(&c->Uy, ne * sizeof(double))) != cudaSuccess) return fail(e, "cudaMalloc Uy");
    if ((e = cudaMalloc(&c->tmp, 5 * ne * sizeof(double))) != cudaSuccess) return fail(e, "cudaMalloc tmp");
    if ((e = cudaMalloc(&c->part, (2 * RED_BLOCKS + 2) * sizeof(double))) != cudaSuccess) return fail(e, "cudaMalloc part");
    if ((e = cudaMallocHost(&c->red_host, 2 * sizeof(double))) != cudaSuccess) return fail(e, "cudaMallocHost");
    {
        std::vector<double> lat(36 * ne, 0.0);
        init_host(c->g, c->p, lat);
        cudaMemcpyAsync(c->lat, lat.data(), 36 * ne * sizeof(double), cudaMemcpyHostToDevice, c->stream);
        cudaMemsetAsync(c->Ux, 0, ne * sizeof(double), c->stream);
        cudaMemsetAsync(c->Uy, 0, ne * sizeof(double), c->stream);
        if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return fail(e, "initial upload");
    }
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return fail(e, "initialisation");
    *out = c;
    return CLBM_OK;
}

int clbm_yl2d_destroy(clbm_yl2d *c)
{
    if (!c) return CLBM_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->lat); cudaFree(c->C); cudaFree(c->Ux); cudaFree(c->Uy); cudaFree(c->tmp); cudaFree(c->part);
    if (c->red_host) cudaFreeHost(c->red_host);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return CLBM_OK;
}

int clbm_yl2d_step(clbm_yl2d *c, int nsteps)
{
    if (!c || nsteps < 0) { set_error("bad argument"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    for (int s = 0; s < nsteps; ++s) {
        int rc = one_step(c);
        if (rc) return rc;
    }
    return CLBM_OK;
}

int clbm_yl2d_step_timed(clbm_yl2d *c, int nsteps, float *ms)
{
    if (!c || nsteps < 0 || !ms) { set_error("bad argument"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    CLBM_CUDA(cudaEventRecord(c->ev0, c->stream));
    int rc = clbm_yl2d_step(c, nsteps);
    if (rc) return rc;
    CLBM_CUDA(cudaEventRecord(c->ev1, c->stream));
    CLBM_CUDA(cudaEventSynchronize(c->ev1));
    CLBM_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return CLBM_OK;
}

int clbm_yl2d_sync(clbm_yl2d *c)
{
    if (!c) { set_error("null context"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    return CLBM_OK;
}

int64_t clbm_yl2d_launch_count(const clbm_yl2d *c) { return c ? c->launches : 0; }

int clbm_yl2d_download_fields(clbm_yl2d *c, double *C, double *P, double *Rho, double *Ux, double *Uy)
{
    if (!c) { set_error("null context"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    int rc = current_fields(c);
    if (rc) return rc;
    const size_t ne = (size_t)c->g.ne, nd = ne * sizeof(double);
    double *host[5] = {C, P, Rho, Ux, Uy};
    for (int j = 0; j < 5; ++j)
        if (host[j]) CLBM_CUDA(cudaMemcpyAsync(host[j], c->tmp + j * ne, nd, cudaMemcpyDeviceToHost, c->stream));
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    return CLBM_OK;
}

int clbm_yl2d_download_lattice(clbm_yl2d *c, double *lattice, int *parity)
{
    if (!c || !lattice) { set_error("bad argument"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    CLBM_CUDA(cudaMemcpyAsync(lattice, c->lat, 36 * (size_t)c->g.ne * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    if (parity) *parity = c->parity;
    return CLBM_OK;
}

int clbm_yl2d_upload(clbm_yl2d *c, const double *lattice, const double *Ux, const double *Uy, int parity)
{
    if (!c || !lattice || !Ux || !Uy || (parity != 0 && parity != 1)) { set_error("bad argument"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    const size_t ne = (size_t)c->g.ne;
    CLBM_CUDA(cudaMemcpyAsync(c->lat, lattice, 36 * ne * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CLBM_CUDA(cudaMemcpyAsync(c->Ux, Ux, ne * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CLBM_CUDA(cudaMemcpyAsync(c->Uy, Uy, ne * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    c->parity = parity;
    c->keep_u = 1;      // the reference's Ux, Uy are the values update_fields already produced for these populations
    return CLBM_OK;
}

int clbm_yl2d_reduce(clbm_yl2d *c, int kind, double *out)
{
    if (!c || !out || (kind != CLBM_REDUCE_MASS && kind != CLBM_REDUCE_ENERGY)) { set_error("bad argument"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    int rc = current_fields(c);
    if (rc) return rc;
    const size_t ne = (size_t)c->g.ne;
    yl2d_reduce1<<<RED_BLOCKS, 256, 0, c->stream>>>(c->tmp + 2 * ne, c->tmp + 3 * ne, c->tmp + 4 * ne, c->g.ne, c->part);
    yl2d_reduce2<<<1, 32, 0, c->stream>>>(c->part, RED_BLOCKS, c->part + 2 * RED_BLOCKS);
    c->launches += 2;
    CLBM_CUDA(cudaMemcpyAsync(c->red_host, c->part + 2 * RED_BLOCKS, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    // totalMass_Young_Laplace2D (AB:436-445) = sum Rho; computeEnergy (AB:425-435) = 0.5 sum(u.u) / (nx ny)
    *out = kind == CLBM_REDUCE_MASS ? c->red_host[0] : 0.5 * c->red_host[1] / ((double)c->g.nx * c->g.ny);
    return CLBM_OK;
}

}  // extern "C"
