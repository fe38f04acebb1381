// hcz3d_kernels.cu -- He-Chen-Zhang phase-field D3Q19 time step (PF/apps/laplace3D.h).
//
// laplace3D.h defines psi_rho = total_P - rho/3 where total_P itself needs grad(lap phi) and
// grad(psi phi) (:318-336), so grad_psi_rho (:470-500) reaches radius 3 in phi; the reference
// recomputes that whole chain at all 19 neighbours (2 kLUPS/core).  Here the levels
// of SURVEY.md A.7 are materialised once per step:
//   level 0  hcz3d_moments_kernel : phi, P_term, raw g momentum            (:216-258)
//   level 1  hcz3d_level1_kernel  : lap phi (wall neighbours skipped, :370-393), psi(phi) (:268-275)
//   level 2  hcz3d_level2_kernel  : u (:280-312 incl. the forcey-in-z quirk), total_P (:318-328),
//                                   psi_rho = total_P - rho/3 (:330-336)
//   level 3  hcz3d_collide_kernel : grad psi_rho (:470-500), collideBgk (:562-624), rest (:664-677),
//                                   push stream (:539-559)
// Wall fallback of the gradients: a bounce_back neighbour contributes the CENTRE value (:450-455).
// Field slots: 0 phi, 1 P_term, 2-4 raw momentum, 5 lap phi, 6 psi(phi), 7 psi_rho
#include <cstdlib>

#include "sc_cell.cuh"

namespace clbm {

using L19 = D3Q19;
CLBM_D double hcz_rho_of_phi3(const ModelParams &mp, double phi)
{
    return mp.rho_g + ((phi - mp.phi_g) / (mp.phi_l - mp.phi_g)) * (mp.rho_l - mp.rho_g);
}

struct FieldPtrs8 { double *p[8]; };

__global__ void __launch_bounds__(256)
hcz3d_moments_kernel(const double *__restrict__ fin, const double *__restrict__ gin, FieldPtrs8 F, Geom g, int x0,
                     long long ncell)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell) return;
    const long long i = (long long)(x0 + g.G) * g.plane + t;
    double f[19];
#pragma unroll
    for (int k = 0; k < 19; ++k) f[k] = fin[(size_t)k * g.ncs + i];
    F.p[0][i] = Mom<L19>::sum(f);
#pragma unroll
    for (int k = 0; k < 19; ++k) f[k] = gin[(size_t)k * g.ncs + i];
    double jx, jy, jz;
    Mom<L19>::first(f, jx, jy, jz);
    F.p[1][i] = Mom<L19>::sum(f);
    F.p[2][i] = jx;
    F.p[3][i] = jy;
    F.p[4][i] = jz;
}

__global__ void __launch_bounds__(256)
hcz3d_level1_kernel(const uint8_t *__restrict__ flag, FieldPtrs8 F, Geom g, ModelParams mp, int x0, long long ncell)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell) return;
    const int x = x0 + (int)(t / g.plane);
    const int r = (int)(t % g.plane);
    const Nbr n = make_nbr(g, x, r / g.nz, r % g.nz);
    const double *__restrict__ phi = F.p[0];
    const double phi_c = phi[n.i];
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < 19; ++k) {
        if (k == L19::REST) continue;
        const long long nb = n.at<L19>(k);
        if (flag[nb] != CELL_BB) sum += L19::t(k) * (phi[nb] - phi_c);
    }
    F.p[5][n.i] = 6.0 * sum;
    F.p[6][n.i] = hcz_psi(phi_c, mp.a, mp.b);
}

struct Grad3 { double x, y, z; };

CLBM_D Grad3 hcz3d_grad(const double *__restrict__ X, const uint8_t *__restrict__ flag, const Nbr &n)
{
    double gx = 0.0, gy = 0.0, gz = 0.0;
    const double xc = X[n.i];
#pragma unroll
    for (int k = 0; k < 19; ++k) {
        if (k == L19::REST) continue;
        const long long nb = n.at<L19>(k);
        const double v = (flag[nb] == CELL_BB) ? xc : X[nb];
        if (L19::cx(k)) gx += L19::t(k) * L19::cx(k) * v;
        if (L19::cy(k)) gy += L19::t(k) * L19::cy(k) * v;
        if (L19::cz(k)) gz += L19::t(k) * L19::cz(k) * v;
    }
    return {3.0 * gx, 3.0 * gy, 3.0 * gz};
}

struct Hcz3dNode {
    double phi, rho, P, u[3];
    Grad3 glap, gpsiphi;
};

CLBM_D void hcz3d_node(const ModelParams &mp, const FieldPtrs8 &F, const uint8_t *flag, const Nbr &n, Hcz3dNode &o)
{
    o.phi = F.p[0][n.i];
    o.rho = hcz_rho_of_phi3(mp, o.phi);
    o.glap = hcz3d_grad(F.p[5], flag, n);
    o.gpsiphi = hcz3d_grad(F.p[6], flag, n);
    const double forcex = mp.kappa * o.phi * o.glap.x;
    double forcey = mp.kappa * o.phi * o.glap.y;
    forcey += mp.gravity * o.rho;
    const double d = o.rho / 3.;
    o.u[0] = (F.p[2][n.i] + forcex / 6.) / d;
    o.u[1] = (F.p[3][n.i] + forcey / 6.) / d;
    o.u[2] = (F.p[4][n.i] + forcey / 6.) / d;  // sic: forcey (laplace3D.h:304)
    o.P = F.p[1][n.i] - 0.5 * (o.u[0] * o.gpsiphi.x + o.u[1] * o.gpsiphi.y + o.u[2] * o.gpsiphi.z);
}

__global__ void __launch_bounds__(256)
hcz3d_level2_kernel(const uint8_t *__restrict__ flag, FieldPtrs8 F, Geom g, ModelParams mp, int x0, long long ncell)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell) return;
    const int x = x0 + (int)(t / g.plane);
    const int r = (int)(t % g.plane);
    const Nbr n = make_nbr(g, x, r / g.nz, r % g.nz);
    Hcz3dNode o;
    hcz3d_node(mp, F, flag, n, o);
    F.p[7][n.i] = o.P - o.rho / 3.0;
}

// MRT = true: CLBM_COLLISION_MRT: out = in + F - M^-1 S M (in - eq + F/2) with the equilibria and forcing terms of collideBgk, the
// forcing without its (1 - omega/2) factor, in the D3Q19 moment basis of mrt.cuh (S = omega I is collideBgk again; the reference has
// no D3Q19 MRT operator: parity unpinned).  Runs on the staged path only (clbm_create clears `fused` for it).
template <bool MRT>
__global__ void __launch_bounds__(256, MRT ? 1 : 2)
hcz3d_collide_kernel(const double *__restrict__ fin, double *__restrict__ fout, const double *__restrict__ gin,
                     double *__restrict__ gout, const uint8_t *__restrict__ flag, FieldPtrs8 F, Geom g,
                     ModelParams mp, int x0, long long ncell)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell) return;
    const int x = x0 + (int)(t / g.plane);
    const int r = (int)(t % g.plane);
    const Nbr n = make_nbr(g, x, r / g.nz, r % g.nz);
    if (flag[n.i] != CELL_BULK) return;

    Hcz3dNode o;
    hcz3d_node(mp, F, flag, n, o);
    const Grad3 E = hcz3d_grad(F.p[7], flag, n);

    const double omega = mp.omega, hw = 1. - 0.5 * omega;
    const double phi = o.phi, rho = o.rho, P = o.P, u0 = o.u[0], u1 = o.u[1], u2 = o.u[2];
    const double forcex = mp.kappa * phi * o.glap.x;
    double forcey = mp.kappa * phi * o.glap.y;
    const double forcez = mp.kappa * phi * o.glap.z;
    forcey += mp.gravity * rho;
    const double usqr = 1.5 * (u0 * u0 + u1 * u1 + u2 * u2);
    const double inv_phi = 1.0 / phi, inv_rho = 1.0 / rho;

    unsigned wall = 0;
#pragma unroll
    for (int k = 0; k < 19; ++k)
        if (k != 9 && flag[n.at<L19>(k)] == CELL_BB) wall |= 1u << k;

    if constexpr (MRT) {
        double Ff[19], Fg[19], vf[19], vg[19], wv[19];
#pragma unroll
        for (int k = 0; k < 19; ++k) {
            const double ck_u = L19::cx(k) * u0 + L19::cy(k) * u1 + L19::cz(k) * u2;
            const double poly = 3 * ck_u + 4.5 * ck_u * ck_u - usqr;
            const double eqf = phi * L19::t(k) * (1 + poly);
            const double eqg = L19::t(k) * (P + (rho / 3.0) * poly);
            const double ex = L19::cx(k) - u0, ey = L19::cy(k) - u1, ez = L19::cz(k) - u2;
            Fg[k] = (ex * forcex + ey * forcey + ez * forcez) * eqf * inv_phi + ((ex * -1 * E.x) + (ey * -1 * E.y) + (ez * -1 * E.z)) * (eqf * inv_phi - L19::t(k));
            Ff[k] = ((ex * -1 * o.gpsiphi.x) + (ey * -1 * o.gpsiphi.y) + (ez * -1 * o.gpsiphi.z)) * 3. * eqf * inv_rho;
            vf[k] = fin[(size_t)k * g.ncs + n.i] - eqf + 0.5 * Ff[k];
            vg[k] = gin[(size_t)k * g.ncs + n.i] - eqg + 0.5 * Fg[k];
        }
        const MrtRates S = {omega, mp.s_e, mp.s_eps, mp.s_q, omega};
        mrt19_relax(vf, S, wv);
#pragma unroll
        for (int k = 0; k < 19; ++k) {
            const double pf = fin[(size_t)k * g.ncs + n.i] + Ff[k] - wv[k];
            if (k == 9) fout[(size_t)9 * g.ncs + n.i] = pf;
            else if (wall & (1u << k)) fout[(size_t)L19::opp(k) * g.ncs + n.i] = pf;
            else fout[(size_t)k * g.ncs + n.at<L19>(k)] = pf;
        }
        mrt19_relax(vg, S, wv);
#pragma unroll
        for (int k = 0; k < 19; ++k) {
            const double pg = gin[(size_t)k * g.ncs + n.i] + Fg[k] - wv[k];
            if (k == 9) gout[(size_t)9 * g.ncs + n.i] = pg;
            else if (wall & (1u << k)) gout[(size_t)L19::opp(k) * g.ncs + n.i] = pg;
            else gout[(size_t)k * g.ncs + n.at<L19>(k)] = pg;
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 19; ++k) {
        const double fk = fin[(size_t)k * g.ncs + n.i];
        const double gk = gin[(size_t)k * g.ncs + n.i];
        double pf, pg;
        if (k == 9) {
            const double eqf0 = phi * L19::t(9) * (1. - usqr);
            const double eqg0 = L19::t(9) * (P - (rho / 3.0) * usqr);
            const double fg0 = hw * -1 * (u0 * forcex + u1 * forcey + u2 * forcez) * eqf0 * inv_phi +
                               hw * -1 * (u0 * -1 * E.x + u1 * -1 * E.y + u2 * -1 * E.z) * (eqf0 * inv_phi - L19::t(9));
            const double ff0 = hw * -3. * eqf0 * (u0 * -1 * o.gpsiphi.x + u1 * -1 * o.gpsiphi.y + u2 * -1 * o.gpsiphi.z) * inv_rho;
            pf = (1 - omega) * fk + omega * eqf0 + ff0;
            pg = (1 - omega) * gk + omega * eqg0 + fg0;
        } else {
            // for k >= 10 the reference derives the equilibrium from the k-10 one (eq - 6 x t ck_u, :574,577);
            // algebraically that is the same polynomial evaluated at c_k
            const double ck_u = L19::cx(k) * u0 + L19::cy(k) * u1 + L19::cz(k) * u2;
            const double poly = 3 * ck_u + 4.5 * ck_u * ck_u - usqr;
            const double eqf = phi * L19::t(k) * (1 + poly);
            const double eqg = L19::t(k) * (P + (rho / 3.0) * poly);
            const double ex = L19::cx(k) - u0, ey = L19::cy(k) - u1, ez = L19::cz(k) - u2;
            const double fg = hw * ((ex * forcex + ey * forcey + ez * forcez) * eqf * inv_phi) +
                              hw * ((ex * -1 * E.x) + (ey * -1 * E.y) + (ez * -1 * E.z)) * (eqf * inv_phi - L19::t(k));
            const double ff = hw * ((ex * -1 * o.gpsiphi.x) + (ey * -1 * o.gpsiphi.y) + (ez * -1 * o.gpsiphi.z)) * 3. * eqf * inv_rho;
            pf = (1. - omega) * fk + omega * eqf + ff;
            pg = (1. - omega) * gk + omega * eqg + fg;
        }
        if (k == 9) {
            fout[(size_t)9 * g.ncs + n.i] = pf;
            gout[(size_t)9 * g.ncs + n.i] = pg;
        } else if (wall & (1u << k)) {
            fout[(size_t)L19::opp(k) * g.ncs + n.i] = pf;
            gout[(size_t)L19::opp(k) * g.ncs + n.i] = pg;
        } else {
            const long long nb = n.at<L19>(k);
            fout[(size_t)k * g.ncs + nb] = pf;
            gout[(size_t)k * g.ncs + nb] = pg;
        }
    }
}

__global__ void __launch_bounds__(256)
hcz3d_fields_kernel(const uint8_t *__restrict__ flag, FieldPtrs8 F, Geom g, ModelParams mp, double *s0, double *s1,
                    double *s2, double *ux, double *uy, double *uz, long long ncell)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell) return;
    const int x = (int)(t / g.plane);
    const int r = (int)(t % g.plane);
    const Nbr n = make_nbr(g, x, r / g.nz, r % g.nz);
    Hcz3dNode o;
    hcz3d_node(mp, F, flag, n, o);
    const bool bulk = flag[n.i] == CELL_BULK;
    if (s0) s0[t] = o.phi;
    if (s1) s1[t] = bulk ? o.P : 0.0;
    if (s2) s2[t] = o.rho;
    if (ux) ux[t] = bulk ? o.u[0] : 0.0;
    if (uy) uy[t] = bulk ? o.u[1] : 0.0;
    if (uz) uz[t] = bulk ? o.u[2] : 0.0;
}

// ---- host side ---------------------------------------------------------------------------
static FieldPtrs8 fld8(clbm_ctx *c)
{
    FieldPtrs8 F;
    for (int i = 0; i < 8; ++i) F.p[i] = c->fld[i];
    return F;
}

int hcz3d_moments(clbm_ctx *c)
{
    c->mom_valid = 0;   // fld[0..4] now hold plain sums: whatever the sweep kernel carried (edge sums, the other set) is stale
    c->sweep_active = 0;
    const long long n = (long long)c->geo.nx * c->geo.plane;
    LaunchScope ls(c, "hcz3d_moments");
    hcz3d_moments_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->pop[1][c->parity], fld8(c), c->geo, 0, n);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

// plain sums of the planes [x0, x0 + np) into moment set `set` of the sweep kernel
static int hcz3d_moments_planes(clbm_ctx *c, int set, int x0, int np)
{
    FieldPtrs8 F = fld8(c);
    for (int m = 0; m < 5; ++m) F.p[m] = c->mom[set][m];
    const long long n = (long long)np * c->geo.plane;
    LaunchScope ls(c, "hcz3d_moments_planes");
    hcz3d_moments_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->pop[1][c->parity], F, c->geo, x0, n);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}
int hcz3d_level1(clbm_ctx *c)
{
    const int x0 = c->multi ? -2 : 0, x1 = c->multi ? c->geo.nx + 2 : c->geo.nx;
    const long long n = (long long)(x1 - x0) * c->geo.plane;
    LaunchScope ls(c, "hcz3d_level1");
    hcz3d_level1_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(c->flag, fld8(c), c->geo, c->mp, x0, n);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}
int hcz3d_level2(clbm_ctx *c)
{
    const int x0 = c->multi ? -1 : 0, x1 = c->multi ? c->geo.nx + 1 : c->geo.nx;
    const long long n = (long long)(x1 - x0) * c->geo.plane;
    LaunchScope ls(c, "hcz3d_level2");
    hcz3d_level2_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(c->flag, fld8(c), c->geo, c->mp, x0, n);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}
int hcz3d_sweep_launch(clbm_ctx *c, int src);    // hcz3d_sweep.cu
bool hcz3d_march_eligible(const clbm_ctx *c);   // hcz3d_march.cu
int hcz3d_march_collide(clbm_ctx *c);

int hcz3d_collide(clbm_ctx *c);
bool hcz3d_fused_eligible(const clbm_ctx *c);   // hcz3d_fused.cu
int hcz3d_fused_launch(clbm_ctx *c, int variant);

// fused == 1 or >= 7: levels 1-3 + collide in one TMA-staged kernel; 2..6: march kernel tile variants
static bool use_fused3d(const clbm_ctx *c)
{
    int v = c->prm.fused;
    if (c->env.hcz_tile >= 0) v = c->env.hcz_tile;
    return (v == 1 || v >= 7) && hcz3d_fused_eligible(c);
}

// everything after the moments: stage 1 of the slab protocol
int hcz3d_stage1(clbm_ctx *c)
{
    if (c->sweep_active) {   // decided in stage 0 of this step
        int rc = hcz3d_sweep_launch(c, c->mom_src);
        if (rc) return rc;
        c->mom_src = 1 - c->mom_src;
        c->mom_valid = 1;
        return 0;
    }
    c->mom_valid = 0;
    if (use_fused3d(c)) {
        int v = c->prm.fused;
        if (c->env.hcz_tile >= 0) v = c->env.hcz_tile;
        return hcz3d_fused_launch(c, v);
    }
    int rc;
    if ((rc = hcz3d_level1(c))) return rc;
    if ((rc = hcz3d_level2(c))) return rc;
    return hcz3d_collide(c);
}

int hcz3d_collide(clbm_ctx *c)
{
    if (c->prm.fused && hcz3d_march_eligible(c)) return hcz3d_march_collide(c);
    const long long n = (long long)c->geo.nx * c->geo.plane;
    LaunchScope ls(c, "hcz3d_collide_stream", true);
    if (c->prm.collision == CLBM_COLLISION_MRT)
        hcz3d_collide_kernel<true><<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->pop[0][1 - c->parity],
                                                                          c->pop[1][c->parity], c->pop[1][1 - c->parity], c->flag,
                                                                          fld8(c), c->geo, c->mp, 0, n);
    else
        hcz3d_collide_kernel<false><<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->pop[0][1 - c->parity],
                                                                           c->pop[1][c->parity], c->pop[1][1 - c->parity], c->flag,
                                                                           fld8(c), c->geo, c->mp, 0, n);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

// ---- single-sweep step (hcz3d_sweep.cu) --------------------------------------------------------------------------
bool hcz3d_sweep_shape_ok(const clbm_ctx *c);
int hcz3d_sweep_launch(clbm_ctx *c, int src);
long long hcz3d_sweep_edge_doubles(const clbm_ctx *c);
int hcz3d_sweep_zero_edges(clbm_ctx *c, int set, int x0, int np);
int hcz3d_sweep_zero_edge_planes(clbm_ctx *c, int set, int xa, int xb);

__global__ void __launch_bounds__(256) count_walls_kernel(const uint8_t *__restrict__ flag, long long n, int *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = (i < n && flag[i] != CELL_BULK) ? 1 : 0;
    if (__syncthreads_or(w) && threadIdx.x == 0) *out = 1;
}

// does the slab hold a bounce_back node?  scanned once per uploaded / initialised state
static int hcz3d_has_walls(clbm_ctx *c, bool *walls)
{
    if (!c->walls_known) {
        const Geom &g = c->geo;
        const long long n = (long long)g.nx * g.plane;
        int *d = reinterpret_cast<int *>(c->red_dev);
        CLBM_CUDA(cudaMemsetAsync(d, 0, sizeof(int), c->stream));
        {
            LaunchScope ls(c, "count_walls");
            count_walls_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(c->flag + (size_t)g.G * g.plane, n, d);
            CLBM_CUDA(cudaGetLastError());
        }
        int *h = reinterpret_cast<int *>(c->red_host);
        CLBM_CUDA(cudaMemcpyAsync(h, d, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CLBM_CUDA(cudaStreamSynchronize(c->stream));
        c->has_walls = *h != 0;
        c->walls_known = 1;
    }
    *walls = c->has_walls != 0;
    return 0;
}

// the default fused path of a single, wall-free slab whose plane is a whole number of 8 x 32 tiles, with enough tiles to fill
// the GPU without x-chunks (CLBM_HCZ3D_SWEEP=1 forces it on small lattices for the tests, =0 turns it off)
static int hcz3d_use_sweep(clbm_ctx *c, bool *use)
{
    *use = false;
    if (c->env.hcz3d_sweep == 0 || c->profiling) return 0;
    int v = c->prm.fused;
    if (c->env.hcz_tile >= 0) v = c->env.hcz_tile;
    if (v != 1 || !hcz3d_fused_eligible(c) || !hcz3d_sweep_shape_ok(c)) return 0;
    const long long tiles = (long long)(c->geo.ny / 8) * (c->geo.nz / 32);
    if (c->env.hcz3d_sweep != 1 && tiles < 128) return 0;
    bool walls = true;
    if (int rc = hcz3d_has_walls(c, &walls)) return rc;
    *use = !walls;
    return 0;
}

static int hcz3d_sweep_alloc(clbm_ctx *c)
{
    if (c->mom[1][0]) return 0;
    const size_t nb = (size_t)c->geo.ncs * sizeof(double), eb = (size_t)hcz3d_sweep_edge_doubles(c) * sizeof(double);
    for (int m = 0; m < 5; ++m) {
        c->mom[0][m] = c->fld[m];
        if (cudaMalloc(&c->mom[1][m], nb) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory (second moment set)"); return CLBM_ENOMEM; }
        CLBM_CUDA(cudaMemsetAsync(c->mom[1][m], 0, nb, c->stream));
    }
    for (int s = 0; s < 2; ++s) {
        // mome[s][0]: edge sums of phi [nx][eplane]; mome[s][1]: of P_term, jx, jy, jz interleaved [nx][eplane][4]
        if (cudaMalloc(&c->mome[s][0], eb) != cudaSuccess || cudaMalloc(&c->mome[s][1], 4 * eb) != cudaSuccess) {
            cudaGetLastError();
            set_error("out of device memory (edge sums)");
            return CLBM_ENOMEM;
        }
        if (int rc = hcz3d_sweep_zero_edges(c, s, 0, c->geo.nx)) return rc;
    }
    return 0;
}

double *hcz3d_moment_array(const clbm_ctx *c, int m) { return (c->sweep_active && c->mom[1][0]) ? c->mom[c->mom_src][m] : c->fld[m]; }

// stage 0 of an x-slab step: complete moments of the local planes + (by the caller) the pack of the moment halo.
// Sweep path: the interior planes come from the previous sweep; planes 0 and nx-1 are rebuilt from the populations, which
// the crossing populations of the neighbours changed after the sweep (their edge sums are therefore zero).
int hcz3d_stage0(clbm_ctx *c, bool rebuild)
{
    int rc;
    bool sweep = false;
    if (rebuild) c->mom_valid = 0;
    if ((rc = hcz3d_use_sweep(c, &sweep))) return rc;
    if (!sweep || rebuild) return hcz3d_moments(c);   // plain sums of every plane into fld[0..4] (= set 0)
    if ((rc = hcz3d_sweep_alloc(c))) return rc;
    const Geom &g = c->geo;
    if (!c->mom_valid) {
        if ((rc = hcz3d_moments(c))) return rc;
        if ((rc = hcz3d_sweep_zero_edges(c, 0, 0, g.nx))) return rc;
        c->mom_src = 0;
    } else {
        if ((rc = hcz3d_moments_planes(c, c->mom_src, 0, 1))) return rc;
        if ((rc = hcz3d_moments_planes(c, c->mom_src, g.nx - 1, 1))) return rc;
        if ((rc = hcz3d_sweep_zero_edge_planes(c, c->mom_src, 0, g.nx - 1))) return rc;
    }
    c->sweep_active = 1;
    return 0;
}

int hcz3d_step(clbm_ctx *c)
{
    int rc;
    bool sweep = false;
    if ((rc = hcz3d_use_sweep(c, &sweep))) return rc;
    if (sweep) {
        if ((rc = hcz3d_sweep_alloc(c))) return rc;
        if (!c->mom_valid) {
            // first step on this state: plain sums into set 0 (= fld[0..4]), whose edge sums are therefore zero
            if ((rc = hcz3d_moments(c))) return rc;
            if ((rc = hcz3d_sweep_zero_edges(c, 0, 0, c->geo.nx))) return rc;
            c->mom_src = 0;
        }
        if ((rc = hcz3d_sweep_launch(c, c->mom_src))) return rc;
        c->mom_src = 1 - c->mom_src;
        c->mom_valid = 1;
        c->parity = 1 - c->parity;
        return 0;
    }
    if ((rc = hcz3d_moments(c))) return rc;
    if ((rc = hcz3d_stage1(c))) return rc;
    c->parity = 1 - c->parity;
    return 0;
}

int hcz3d_fields(clbm_ctx *c, double *s0, double *s1, double *s2, double *ux, double *uy, double *uz)
{
    int rc;
    if (!c->multi) {
        if ((rc = hcz3d_moments(c))) return rc;
    }
    if ((rc = hcz3d_level1(c))) return rc;
    const long long n = (long long)c->geo.nx * c->geo.plane;
    LaunchScope ls(c, "hcz3d_fields");
    hcz3d_fields_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(c->flag, fld8(c), c->geo, c->mp, s0, s1, s2, ux, uy, uz, n);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace clbm
