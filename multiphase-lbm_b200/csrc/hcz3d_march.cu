// hcz3d_march.cu -- plane-marching collide/stream kernel of the HCZ D3Q19 model (PF/apps/laplace3D.h).
//
// The staged collide kernel (hcz3d_kernels.cu) fetches three 18-point stencils (lap phi, psi(phi), psi_rho)
// plus 18 mask bytes per cell through L1.  Here a CTA owns a TY x TZ (y,z) tile and marches along x; the
// three stencil fields and the node mask of plane x+1 (tile + halo ring) are staged into 4-slot shared-memory
// rings, so every field value is fetched once per tile and the gradients read shared memory.  The 2 x 19
// populations of a cell stream through registers exactly once (loaded, relaxed with both HCZ forcing terms,
// pushed to the neighbour with half-way bounce-back).
//
//   grad X            laplace3D.h:435-536  (bounce_back neighbour -> centre value)
//   velocity/total_P  laplace3D.h:280-328  (incl. the forcey-in-z quirk)
//   collideBgk / rest laplace3D.h:562-624, :664-677 ;  stream :539-559
// Field slots (see hcz3d_kernels.cu): 0 phi, 1 P_term, 2-4 raw momentum, 5 lap phi, 6 psi(phi), 7 psi_rho.
#include <cstdlib>

#include "sc_cell.cuh"

namespace clbm {

using L19m = D3Q19;

struct HczTables {
    const double *fin[19];
    const double *gin[19];
    double *fout[19];
    double *gout[19];
};
struct HczFields { const double *p[8]; };

template <int TY, int TZ, int MINB>
__global__ void __launch_bounds__(TY *TZ, MINB)
hcz3d_march_kernel(const HczTables P, const HczFields F, const uint8_t *__restrict__ flag, Geom g, ModelParams mp, int xchunk)
{
    constexpr int NT = TY * TZ, SY = TY + 2, SZ = TZ + 2;
    __shared__ double r_lap[4][SY][SZ], r_pp[4][SY][SZ], r_pr[4][SY][SZ];
    __shared__ uint8_t r_fl[4][SY][SZ];

    const int tid = threadIdx.x;
    const int tz = tid % TZ, ty = tid / TZ;
    const int y0 = blockIdx.y * TY, z0 = blockIdx.x * TZ;
    const int y = y0 + ty, z = z0 + tz;
    const bool inside = (y < g.ny) && (z < g.nz);
    const int xa = blockIdx.z * xchunk;
    const int xb = min(g.nx, xa + xchunk);
    const int plane = (int)g.plane, nz = g.nz, G = g.G;
    const int ty_n = min(TY, g.ny - y0), tz_n = min(TZ, g.nz - z0);
    const int nrow = tz_n + 2, nhalo = 2 * nrow + 2 * ty_n;
    const int yz = y * nz + z;

    // stage lap phi, psi(phi), psi_rho and the mask of storage plane xs into ring slot `slot`
    auto fill = [&](int xs, int slot) {
        if (inside) {
            const int i = xs * plane + yz;
            r_lap[slot][ty + 1][tz + 1] = F.p[5][i];
            r_pp[slot][ty + 1][tz + 1] = F.p[6][i];
            r_pr[slot][ty + 1][tz + 1] = F.p[7][i];
            r_fl[slot][ty + 1][tz + 1] = flag[i];
        }
        for (int h = tid; h < nhalo; h += NT) {
            int sy, sz;
            if (h < nrow) { sy = 0; sz = h; }
            else if (h < 2 * nrow) { sy = ty_n + 1; sz = h - nrow; }
            else { const int q = h - 2 * nrow; sy = 1 + (q >> 1); sz = (q & 1) ? tz_n + 1 : 0; }
            const int i = xs * plane + g.wy(y0 + sy - 1) * nz + g.wz(z0 + sz - 1);
            r_lap[slot][sy][sz] = F.p[5][i];
            r_pp[slot][sy][sz] = F.p[6][i];
            r_pr[slot][sy][sz] = F.p[7][i];
            r_fl[slot][sy][sz] = flag[i];
        }
    };

    fill(g.wx(xa - 1) + G, (xa + 3) & 3);
    fill(xa + G, xa & 3);

    const int oym = (g.wy(y - 1) - y) * nz, oyp = (g.wy(y + 1) - y) * nz;
    const int ozm = g.wz(z - 1) - z, ozp = g.wz(z + 1) - z;
    const double omega = mp.omega, om1 = 1. - omega, hw = 1. - 0.5 * omega;

    for (int x = xa; x < xb; ++x) {
        const int xp = g.wx(x + 1), xm = g.wx(x - 1);
        fill(xp + G, (x + 1) & 3);
        __syncthreads();
        const int sm = (x + 3) & 3, s0 = x & 3, sp = (x + 1) & 3;
        if (!inside || r_fl[s0][ty + 1][tz + 1] != CELL_BULK) continue;

        const int i = (x + G) * plane + yz;
        // gradients of the three staged fields; a bounce_back neighbour contributes the centre value
        unsigned wall = 0;
        double gl[3] = {0., 0., 0.}, gp[3] = {0., 0., 0.}, ge[3] = {0., 0., 0.};
        {
            const double c_lap = r_lap[s0][ty + 1][tz + 1], c_pp = r_pp[s0][ty + 1][tz + 1], c_pr = r_pr[s0][ty + 1][tz + 1];
#pragma unroll
            for (int k = 0; k < 19; ++k) {
                if (k == 9) continue;
                const int slot = L19m::cx(k) < 0 ? sm : (L19m::cx(k) > 0 ? sp : s0);
                const int sy = ty + 1 + L19m::cy(k), sz = tz + 1 + L19m::cz(k);
                const bool w = r_fl[slot][sy][sz] == CELL_BB;
                if (w) wall |= 1u << k;
                const double vl = w ? c_lap : r_lap[slot][sy][sz];
                const double vp = w ? c_pp : r_pp[slot][sy][sz];
                const double ve = w ? c_pr : r_pr[slot][sy][sz];
                const double t = L19m::t(k);
                if (L19m::cx(k)) { gl[0] += t * L19m::cx(k) * vl; gp[0] += t * L19m::cx(k) * vp; ge[0] += t * L19m::cx(k) * ve; }
                if (L19m::cy(k)) { gl[1] += t * L19m::cy(k) * vl; gp[1] += t * L19m::cy(k) * vp; ge[1] += t * L19m::cy(k) * ve; }
                if (L19m::cz(k)) { gl[2] += t * L19m::cz(k) * vl; gp[2] += t * L19m::cz(k) * vp; ge[2] += t * L19m::cz(k) * ve; }
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) { gl[d] *= 3.0; gp[d] *= 3.0; ge[d] *= 3.0; }
        }
        const double phi = F.p[0][i];
        const double rho = mp.rho_g + ((phi - mp.phi_g) / (mp.phi_l - mp.phi_g)) * (mp.rho_l - mp.rho_g);
        const double Fx = mp.kappa * phi * gl[0];
        const double Fy = mp.kappa * phi * gl[1] + mp.gravity * rho;
        const double Fz = mp.kappa * phi * gl[2];
        const double inv_d = 3.0 / rho;          // 1 / (rho/3)
        const double u0 = (F.p[2][i] + Fx / 6.) * inv_d;
        const double u1 = (F.p[3][i] + Fy / 6.) * inv_d;
        const double u2 = (F.p[4][i] + Fy / 6.) * inv_d;   // sic: forcey (laplace3D.h:304)
        const double Pt = F.p[1][i] - 0.5 * (u0 * gp[0] + u1 * gp[1] + u2 * gp[2]);
        const double usqr = 1.5 * (u0 * u0 + u1 * u1 + u2 * u2);
        // (e_k - u).V = c_k.V - u.V for the three forcing vectors V = F, -E, -grad psi(phi)
        const double uF = u0 * Fx + u1 * Fy + u2 * Fz;
        const double uE = u0 * ge[0] + u1 * ge[1] + u2 * ge[2];
        const double uG = u0 * gp[0] + u1 * gp[1] + u2 * gp[2];
        const double rho3 = rho / 3.0;
        const double ffs = hw * 3.0 * phi / rho;   // ff = hw * C * 3 * eqf / rho,  eqf = phi * Gamma
        const int oxm = (xm - x) * plane, oxp = (xp - x) * plane;

#pragma unroll
        for (int k = 0; k < 19; ++k) {
            const double fk = P.fin[k][i], gk = P.gin[k][i];
            const double t = L19m::t(k);
            double pf, pg;
            if (k == 9) {
                const double Gam = t * (1. - usqr);                   // eqf0 / phi
                const double eqg0 = t * (Pt - rho3 * usqr);
                // fg0 = hw*(-(u.F) Gam) + hw*(-(u.(-E))) (Gam - t) ; ff0 = hw*(-3) eqf0 (u.(-G)) / rho
                const double fg0 = hw * (-uF * Gam + uE * (Gam - t));
                const double ff0 = ffs * uG * Gam;
                pf = om1 * fk + omega * phi * Gam + ff0;
                pg = om1 * gk + omega * eqg0 + fg0;
            } else {
                const double cu = L19m::cx(k) * u0 + L19m::cy(k) * u1 + L19m::cz(k) * u2;
                const double poly = 3. * cu + 4.5 * cu * cu - usqr;
                const double Gam = t * (1. + poly);                   // eqf / phi
                const double eqg = t * (Pt + rho3 * poly);
                const double cF = L19m::cx(k) * Fx + L19m::cy(k) * Fy + L19m::cz(k) * Fz;
                const double cE = L19m::cx(k) * ge[0] + L19m::cy(k) * ge[1] + L19m::cz(k) * ge[2];
                const double cG = L19m::cx(k) * gp[0] + L19m::cy(k) * gp[1] + L19m::cz(k) * gp[2];
                const double fg = hw * ((cF - uF) * Gam - (cE - uE) * (Gam - t));
                const double ff = -ffs * (cG - uG) * Gam;
                pf = om1 * fk + omega * phi * Gam + ff;
                pg = om1 * gk + omega * eqg + fg;
            }
            if (k == 9) { P.fout[k][i] = pf; P.gout[k][i] = pg; continue; }
            const int off = (L19m::cx(k) < 0 ? oxm : (L19m::cx(k) > 0 ? oxp : 0)) + (L19m::cy(k) < 0 ? oym : (L19m::cy(k) > 0 ? oyp : 0)) +
                            (L19m::cz(k) < 0 ? ozm : (L19m::cz(k) > 0 ? ozp : 0));
            if (wall & (1u << k)) { P.fout[L19m::opp(k)][i] = pf; P.gout[L19m::opp(k)][i] = pg; }
            else { P.fout[k][i + off] = pf; P.gout[k][i + off] = pg; }
        }
    }
}

template <int TY, int TZ, int MINB>
static int launch_march(clbm_ctx *c)
{
    const Geom &g = c->geo;
    const int tiles = ((g.ny + TY - 1) / TY) * ((g.nz + TZ - 1) / TZ);
    int xchunk = g.nx;
    const long long want = 8LL * 148 * MINB;
    if ((long long)tiles < want) {
        const long long nch = (want + tiles - 1) / tiles;
        xchunk = (int)((g.nx + nch - 1) / nch);
        if (xchunk < 8) xchunk = g.nx < 8 ? g.nx : 8;
    }
    if (c->env.hcz_xchunk > 0) xchunk = c->env.hcz_xchunk < g.nx ? c->env.hcz_xchunk : g.nx;
    dim3 grid((g.nz + TZ - 1) / TZ, (g.ny + TY - 1) / TY, (g.nx + xchunk - 1) / xchunk);
    HczTables P;
    for (int k = 0; k < 19; ++k) {
        P.fin[k] = c->pop[0][c->parity] + (size_t)k * g.ncs;
        P.gin[k] = c->pop[1][c->parity] + (size_t)k * g.ncs;
        P.fout[k] = c->pop[0][1 - c->parity] + (size_t)k * g.ncs;
        P.gout[k] = c->pop[1][1 - c->parity] + (size_t)k * g.ncs;
    }
    HczFields F;
    for (int j = 0; j < 8; ++j) F.p[j] = c->fld[j];
    LaunchScope ls(c, "hcz3d_march_collide_stream", true);
    hcz3d_march_kernel<TY, TZ, MINB><<<grid, TY * TZ, 0, c->stream>>>(P, F, c->flag, g, c->mp, xchunk);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

bool hcz3d_march_eligible(const clbm_ctx *c) { return c->geo.ncs < (1LL << 31); }

int hcz3d_march_collide(clbm_ctx *c)
{
    int variant = c->prm.fused > 1 ? c->prm.fused : 0;
    if (c->env.hcz_tile >= 0) variant = c->env.hcz_tile;
    switch (variant) {
    case 2: return launch_march<8, 32, 2>(c);
    case 3: return launch_march<4, 64, 1>(c);
    case 4: return launch_march<2, 64, 4>(c);
    case 5: return launch_march<4, 32, 4>(c);
    case 6: return launch_march<16, 16, 2>(c);
    default: return launch_march<4, 64, 2>(c);
    }
}

}  // namespace clbm
