// hcz2d_kernels.cu -- He-Chen-Zhang phase-field D2Q9 time step (PF/apps/rayleighTaylor2D.h).
//
// The reference functor re-derives every field recursively per neighbour (~10^4 cached loads
// per cell).  Here the dependency levels of SURVEY.md A.7 are materialised once:
//   level 0  hcz2d_phi_kernel    : phi = sum_k f_k                               (macro_phi_P :197-214)
//   level 1  hcz2d_level1_kernel : lap(phi) with the wall mirror rule (:467-495),
//                                  psi(phi) (:237-242), psi(rho) (:374-379), rho(phi) (:232-235)
//   level 2  hcz2d_collide_kernel: grad lap phi, grad psi_phi, grad psi_rho, grad rho (:341-446,:501-529),
//                                  velocity (:316-337), total_P (:452-460), collideBgk (:552-606),
//                                  rest population (:642-663), push stream with bounce-back (:533-549)
// Field slots: fld[0]=phi  fld[1]=lap phi  fld[2]=psi(phi)  fld[3]=psi(rho)  fld[4]=rho
#include "mrt.cuh"
#include "sc_cell.cuh"

namespace clbm {

using L9 = D2Q9;

CLBM_D double hcz_rho_of_phi(const ModelParams &mp, double phi)
{
    return mp.rho_g + ((phi - mp.phi_g) / (mp.phi_l - mp.phi_g)) * (mp.rho_l - mp.rho_g);
}

__global__ void __launch_bounds__(256)
hcz2d_phi_kernel(const double *__restrict__ fin, double *__restrict__ phi, Geom g, int x0, long long ncell)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell) return;
    const long long i = (long long)(x0 + g.G) * g.plane + t;
    double f[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) f[k] = fin[(size_t)k * g.ncs + i];
    phi[i] = Mom<L9>::sum(f);
}

// value of field X at the k-th neighbour with the mirror rule: a bounce_back neighbour is
// replaced by the opposite neighbour i - c_k (PF/apps/rayleighTaylor2D.h:261-269)
CLBM_D double mirror_at(int k, const double *__restrict__ X, const uint8_t *__restrict__ flag, const Nbr &n)
{
    const long long nb = n.at<L9>(k);
    if (flag[nb] == CELL_BB) return X[n.at<L9>(L9::opp(k))];
    return X[nb];
}

__global__ void __launch_bounds__(256)
hcz2d_level1_kernel(const double *__restrict__ phi, const uint8_t *__restrict__ flag, double *__restrict__ lap,
                    double *__restrict__ psiphi, double *__restrict__ psirho, double *__restrict__ rhoa,
                    Geom g, ModelParams mp, int x0, long long ncell)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell) return;
    const int x = x0 + (int)(t / g.plane);
    const int y = (int)(t % g.plane);
    const Nbr n = make_nbr(g, x, y, 0);
    const double phi_c = phi[n.i];
    const double rho = hcz_rho_of_phi(mp, phi_c);
    rhoa[n.i] = rho;
    psiphi[n.i] = hcz_psi(phi_c, mp.a, mp.b);
    psirho[n.i] = hcz_psi(rho, mp.a, mp.b);
    double sum = 0.0;
    if (flag[n.i] == CELL_BULK) {
#pragma unroll
        for (int k = 0; k < 9; ++k)
            if (k != L9::REST) sum += L9::t(k) * (mirror_at(k, phi, flag, n) - phi_c);
    }
    lap[n.i] = 6.0 * sum;
}

struct Grad2 { double x, y; };

// 3 * sum_k t_k c_k X(nb or mirror)
CLBM_D Grad2 hcz2d_grad(const double *__restrict__ X, const uint8_t *__restrict__ flag, const Nbr &n)
{
    double gx = 0.0, gy = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (k == L9::REST) continue;
        const double v = mirror_at(k, X, flag, n);
        if (L9::cx(k)) gx += L9::t(k) * L9::cx(k) * v;
        if (L9::cy(k)) gy += L9::t(k) * L9::cy(k) * v;
    }
    return {3.0 * gx, 3.0 * gy};
}

struct Hcz2dNode {
    double phi, rho, P, ux, uy;
    Grad2 glap, gpsiphi, gpsirho;
};

// velocity (:316-337) and total_P (:452-460) of one bulk node
CLBM_D void hcz2d_node(const ModelParams &mp, const double *g9, const double *const *fld, const uint8_t *flag,
                       const Nbr &n, Hcz2dNode &o)
{
    o.phi = fld[0][n.i];
    o.rho = fld[4][n.i];
    o.glap = hcz2d_grad(fld[1], flag, n);
    double jx, jy, jz;
    Mom<L9>::first(g9, jx, jy, jz);
    const double Pt = Mom<L9>::sum(g9);
    double forcex, forcey;
    hcz2d_force(mp, o.rho, o.glap.x, o.glap.y, forcex, forcey);
    o.ux = (jx + forcex / 6.0) / (o.rho / 3.0);
    o.uy = (jy + forcey / 6.0) / (o.rho / 3.0);
    const Grad2 grho = hcz2d_grad(fld[4], flag, n);
    o.P = Pt - 0.5 * (o.ux * -grho.x / 3. + o.uy * -grho.y / 3.);
}

struct FieldPtrs5 { const double *p[5]; };

// MRT = true: CLBM_COLLISION_MRT (include/clbm.h): out = in + F - M^-1 S M (in - eq + F/2) with the equilibria and the
// forcing terms of collideBgk, the forcing without its (1 - omega/2) factor (mrt.cuh; S = omega I is collideBgk again)
template <bool MRT>
__global__ void __launch_bounds__(256, 2)
hcz2d_collide_kernel(const double *__restrict__ fin, double *__restrict__ fout, const double *__restrict__ gin,
                     double *__restrict__ gout, const uint8_t *__restrict__ flag, FieldPtrs5 F,
                     Geom g, ModelParams mp, int x0, long long ncell)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell) return;
    const int x = x0 + (int)(t / g.plane);
    const int y = (int)(t % g.plane);
    const Nbr n = make_nbr(g, x, y, 0);
    if (flag[n.i] != CELL_BULK) return;

    double f[9], gg[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        f[k] = fin[(size_t)k * g.ncs + n.i];
        gg[k] = gin[(size_t)k * g.ncs + n.i];
    }
    Hcz2dNode o;
    hcz2d_node(mp, gg, F.p, flag, n, o);
    o.gpsirho = hcz2d_grad(F.p[3], flag, n);
    o.gpsiphi = hcz2d_grad(F.p[2], flag, n);

    const double omega = mp.omega, hw = 1. - 0.5 * omega;
    const double u0 = o.ux, u1 = o.uy, phi = o.phi, rho = o.rho, P = o.P;
    const double usqr = 1.5 * (u0 * u0 + u1 * u1);
    double forcex, forcey, forcex0, forcey0;
    hcz2d_force(mp, rho, o.glap.x, o.glap.y, forcex, forcey);
    // the layered variant drives its REST population with grad lap rho (PF/apps/twoLayeredFlow2D.h:595-598);
    // rho is an affine function of phi, so grad lap rho = (rho_l - rho_g)/(phi_l - phi_g) * grad lap phi, mirror rule included
    const double slope = (mp.sc_force == CLBM_HCZ_FORCE_LAYERED) ? mp.drho * mp.inv_dphi : 1.0;
    hcz2d_force(mp, rho, slope * o.glap.x, slope * o.glap.y, forcex0, forcey0);
    const double Ex = o.gpsirho.x, Ey = o.gpsirho.y;
    const double inv_phi = 1.0 / phi;

    unsigned wall = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k)
        if (k != 4 && flag[n.at<L9>(k)] == CELL_BB) wall |= 1u << k;

    double pf[9], pg[9];
    if constexpr (MRT) {
        double Ff[9], Fg[9], vf[9], vg[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const double ck_u = L9::cx(k) * u0 + L9::cy(k) * u1;
            const double poly = 3 * ck_u + 4.5 * ck_u * ck_u - usqr;      // k = 4: -usqr
            const double eqf = phi * L9::t(k) * (1 + poly);
            const double eqg = L9::t(k) * (P + (rho / 3.0) * poly);
            const double e_u_x = L9::cx(k) - u0, e_u_y = L9::cy(k) - u1;
            if (k == 4) {   // rest population: the layered variant's own force and the reference's (u.(-E)) sign
                Fg[k] = -(u0 * forcex0 + u1 * forcey0) * eqf * inv_phi + ((u0 * -Ex + u1 * -Ey) * (eqf * inv_phi - L9::t(4)));
            } else {
                Fg[k] = (e_u_x * forcex + e_u_y * forcey) * eqf * inv_phi + ((e_u_x * -Ex) + (e_u_y * -Ey)) * (eqf * inv_phi - L9::t(k));
            }
            Ff[k] = ((e_u_x * -o.gpsiphi.x) + (e_u_y * -o.gpsiphi.y)) * 3.0 * eqf * inv_phi;
            vf[k] = f[k] - eqf + 0.5 * Ff[k];
            vg[k] = gg[k] - eqg + 0.5 * Fg[k];
        }
        const MrtRates S = {omega, mp.s_e, mp.s_eps, mp.s_q, omega};
        double wf[9], wg[9];
        mrt9_relax(vf, S, wf);
        mrt9_relax(vg, S, wg);
#pragma unroll
        for (int k = 0; k < 9; ++k) { pf[k] = f[k] + Ff[k] - wf[k]; pg[k] = gg[k] + Fg[k] - wg[k]; }
    } else {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (k == 4) {
            const double eqf0 = phi * L9::t(4) * (1. - usqr);
            const double eqg0 = L9::t(4) * (P - (rho / 3.0) * usqr);
            const double fg0 = hw * (-(u0 * forcex0 + u1 * forcey0) * eqf0 * inv_phi +
                                     ((u0 * -Ex + u1 * -Ey) * (eqf0 * inv_phi - L9::t(4))));
            const double ff0 = hw * (-3.0 * (u0 * -o.gpsiphi.x + u1 * -o.gpsiphi.y) * eqf0 * inv_phi);
            pf[k] = (1 - omega) * f[4] + omega * eqf0 + ff0;
            pg[k] = (1 - omega) * gg[4] + omega * eqg0 + fg0;
        } else {
            const double ck_u = L9::cx(k) * u0 + L9::cy(k) * u1;
            const double poly = 3 * ck_u + 4.5 * ck_u * ck_u - usqr;
            const double eqf = phi * L9::t(k) * (1 + poly);
            const double eqg = L9::t(k) * (P + (rho / 3.0) * poly);
            const double e_u_x = L9::cx(k) - u0, e_u_y = L9::cy(k) - u1;
            // note: the reference uses t[k] (not t[opp k]) in the opposite-direction term (:591); equal by symmetry
            const double fg = hw * ((e_u_x * forcex + e_u_y * forcey) * eqf * inv_phi) +
                              hw * ((e_u_x * -Ex) + (e_u_y * -Ey)) * (eqf * inv_phi - L9::t(k));
            const double ff = hw * ((e_u_x * -o.gpsiphi.x) + (e_u_y * -o.gpsiphi.y)) * 3.0 * eqf * inv_phi;
            pf[k] = (1. - omega) * f[k] + omega * eqf + ff;
            pg[k] = (1. - omega) * gg[k] + omega * eqg + fg;
        }
    }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (k == 4) {
            fout[(size_t)4 * g.ncs + n.i] = pf[k];
            gout[(size_t)4 * g.ncs + n.i] = pg[k];
        } else if (wall & (1u << k)) {
            fout[(size_t)L9::opp(k) * g.ncs + n.i] = pf[k];
            gout[(size_t)L9::opp(k) * g.ncs + n.i] = pg[k];
        } else {
            const long long nb = n.at<L9>(k);
            fout[(size_t)k * g.ncs + nb] = pf[k];
            gout[(size_t)k * g.ncs + nb] = pg[k];
        }
    }
}

__global__ void __launch_bounds__(256)
hcz2d_fields_kernel(const double *__restrict__ gin, const uint8_t *__restrict__ flag, FieldPtrs5 F, Geom g,
                    ModelParams mp, double *s0, double *s1, double *s2, double *ux, double *uy, double *uz,
                    long long ncell)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell) return;
    const int x = (int)(t / g.plane);
    const int y = (int)(t % g.plane);
    const Nbr n = make_nbr(g, x, y, 0);
    double P = 0.0, u0 = 0.0, u1 = 0.0;
    if (flag[n.i] == CELL_BULK) {
        double gg[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) gg[k] = gin[(size_t)k * g.ncs + n.i];
        Hcz2dNode o;
        hcz2d_node(mp, gg, F.p, flag, n, o);
        P = o.P; u0 = o.ux; u1 = o.uy;
    }
    if (s0) s0[t] = F.p[0][n.i];
    if (s1) s1[t] = P;
    if (s2) s2[t] = F.p[4][n.i];
    if (ux) ux[t] = u0;
    if (uy) uy[t] = u1;
    if (uz) uz[t] = 0.0;
}

// ---- host side ---------------------------------------------------------------------------
static FieldPtrs5 fld5(clbm_ctx *c)
{
    FieldPtrs5 F;
    for (int i = 0; i < 5; ++i) F.p[i] = c->fld[i];
    return F;
}

int hcz2d_phi(clbm_ctx *c)
{
    const long long n = (long long)c->geo.nx * c->geo.plane;
    LaunchScope ls(c, "hcz2d_phi");
    hcz2d_phi_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->fld[0], c->geo, 0, n);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

int hcz2d_level1(clbm_ctx *c)
{
    // slab mode: also the first ghost plane on each side (phi ghosts reach depth 2)
    const int x0 = c->multi ? -1 : 0, x1 = c->multi ? c->geo.nx + 1 : c->geo.nx;
    const long long n = (long long)(x1 - x0) * c->geo.plane;
    LaunchScope ls(c, "hcz2d_level1");
    hcz2d_level1_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(c->fld[0], c->flag, c->fld[1], c->fld[2], c->fld[3],
                                                               c->fld[4], c->geo, c->mp, x0, n);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

int hcz2d_collide(clbm_ctx *c)
{
    const long long n = (long long)c->geo.nx * c->geo.plane;
    LaunchScope ls(c, "hcz2d_collide_stream", true);
    if (c->prm.collision == CLBM_COLLISION_MRT)
        hcz2d_collide_kernel<true><<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->pop[0][1 - c->parity],
                                                                          c->pop[1][c->parity], c->pop[1][1 - c->parity], c->flag,
                                                                          fld5(c), c->geo, c->mp, 0, n);
    else
        hcz2d_collide_kernel<false><<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->pop[0][1 - c->parity],
                                                                           c->pop[1][c->parity], c->pop[1][1 - c->parity], c->flag,
                                                                           fld5(c), c->geo, c->mp, 0, n);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

bool hcz2d_fused_eligible(const clbm_ctx *c);   // hcz2d_fused.cu
int hcz2d_fused_launch(clbm_ctx *c);

// slab protocol halves (clbm_step_stage): stage 0 makes the phi planes the neighbours need, stage 1 collides
int hcz2d_stage0(clbm_ctx *c)
{
    const Geom &g = c->geo;
    if (!(c->prm.fused && hcz2d_fused_eligible(c)) || g.nx < 4) return hcz2d_phi(c);
    // fused path: only the two boundary columns on each side leave the SMs (the moment halo of depth 2)
    for (int side = 0; side < 2; ++side) {
        const int x0 = side ? g.nx - 2 : 0;
        LaunchScope ls(c, "hcz2d_phi_boundary");
        hcz2d_phi_kernel<<<grid_for(2 * g.plane, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->fld[0], g, x0, 2 * g.plane);
        CLBM_CUDA(cudaGetLastError());
    }
    return 0;
}
int hcz2d_stage1(clbm_ctx *c)
{
    if (c->prm.fused && hcz2d_fused_eligible(c)) return hcz2d_fused_launch(c);
    int rc = hcz2d_level1(c);
    return rc ? rc : hcz2d_collide(c);
}

int hcz2d_step(clbm_ctx *c)
{
    int rc;
    if (c->prm.fused && hcz2d_fused_eligible(c)) {
        if ((rc = hcz2d_fused_launch(c))) return rc;
        c->parity = 1 - c->parity;
        return 0;
    }
    if ((rc = hcz2d_phi(c))) return rc;
    if ((rc = hcz2d_level1(c))) return rc;
    if ((rc = hcz2d_collide(c))) return rc;
    c->parity = 1 - c->parity;
    return 0;
}

int hcz2d_fields(clbm_ctx *c, double *s0, double *s1, double *s2, double *ux, double *uy, double *uz)
{
    int rc;
    // phi of the local columns; in slab mode the caller ran stage 0 + exchange, so the phi ghosts are valid already
    if ((rc = hcz2d_phi(c))) return rc;
    if ((rc = hcz2d_level1(c))) return rc;
    const long long n = (long long)c->geo.nx * c->geo.plane;
    LaunchScope ls(c, "hcz2d_fields");
    hcz2d_fields_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[1][c->parity], c->flag, fld5(c), c->geo, c->mp,
                                                               s0, s1, s2, ux, uy, uz, n);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace clbm
