// hcz3d_fused.cu -- HCZ D3Q19 (PF/apps/laplace3D.h): levels 1-3 and the collide/stream sweep in ONE plane-marching
// kernel with TMA-staged populations (sm_100a).
//
// The step is two launches: hcz3d_moments_kernel (phi, P_term, raw g momentum: 304 B read + 40 B written per
// node) and this kernel.  A CTA owns a TY x TZ (y,z) tile and marches along x.  Per iteration (collide plane x):
//
//   S1  phi, node mask of plane x+3 on the tile + halo 3          -> r_phi, r_fl rings   (prefetched one plane ahead)
//   S2  lap(phi) (:370-393, wall neighbours skipped) and psi(phi) (:268-275) of plane x+2 on tile + halo 2
//   S3  plane x+1 on tile + halo 1: grad lap(phi), grad psi(phi) (:435-536, wall -> centre value), velocity with the
//       forcey-in-z quirk (:280-312), total_P (:318-328), psi_rho = total_P - rho/3 (:330-336)   -> r_pr ring;
//       the owning thread keeps u, P, F, grad psi(phi) of its cell in registers for the collision one plane later
//   S4  plane x: grad psi_rho from the ring, collideBgk of the 2 x 19 populations (:562-624, rest :664-677),
//       push stream with half-way bounce-back (:539-559)
//
// The 38 populations of the tile's plane arrive by two cp.async.bulk.tensor (TMA) boxes [19][TY][TZ] per plane into a
// two-stage shared-memory pipeline (mbarrier expect_tx/complete_tx), issued two planes ahead, so HBM latency is off
// the instruction stream and no scalar field other than the five moments round-trips through HBM.
// Traffic per node: 304 (moments pass) + ~45 (moment fields incl. halo re-reads, L2) + 608 (populations in/out).
#include <cstdlib>
#include <type_traits>

#include "sc_cell.cuh"
#include "tma.cuh"

namespace clbm {

using L19f = D3Q19;

struct Hcz3dOut {
    double *fout[19];
    double *gout[19];
};
struct Hcz3dMom { const double *phi, *pt, *jx, *jy, *jz; };

template <int TY, int TZ>
struct Hcz3dCfg {
    static constexpr int NT = TY * TZ;
    static constexpr int Y3 = TY + 6, Z3 = TZ + 6, R3 = Y3 * Z3;   // tile + halo 3
    static constexpr int Y2 = TY + 4, Z2 = TZ + 4, R2 = Y2 * Z2;   // tile + halo 2
    static constexpr int Y1 = TY + 2, Z1 = TZ + 2;                 // tile + halo 1
    static constexpr int NH1 = 2 * Z1 + 2 * TY;                    // cells of the halo-1 ring
    static constexpr int N3 = (R3 + NT - 1) / NT;                  // S1 cells per thread
    static constexpr int N2 = (R2 + NT - 1) / NT;                  // S2 cells per thread
    static constexpr int SET_BYTES = 19 * NT * 8;
    static constexpr int STAGE_BYTES = 2 * SET_BYTES;
    static constexpr int OFF_PHI = 2 * STAGE_BYTES;
    static constexpr int OFF_LAP = OFF_PHI + 4 * R3 * 8;
    static constexpr int OFF_PP = OFF_LAP + 4 * R2 * 8;
    static constexpr int OFF_PR = OFF_PP + 4 * R2 * 8;
    static constexpr int OFF_FL = OFF_PR + 4 * Y1 * Z1 * 8;
    static constexpr int OFF_BAR = ((OFF_FL + 8 * R3 + 15) / 16) * 16;
    static constexpr int SMEM = OFF_BAR + 32;
    static_assert(NH1 <= NT, "one halo-1 cell per thread");
    static_assert(TZ % 2 == 0, "TMA rows must be a multiple of 16 bytes");
    static_assert(SET_BYTES % 128 == 0, "stage alignment");
};

// 3 * sum_k t_k c_k X(nb), a bounce_back neighbour contributing the centre value; R is a ring of planes with row
// length ZR, (a, b) the node's position in it
template <int ZR, bool WALLS>
CLBM_D void grad19(const double *Rm, const double *R0, const double *Rp, int q, unsigned wall, double g[3])
{
    const double xc = WALLS ? R0[q] : 0.0;
    // opposite directions are paired first (the weights are equal), then the pair differences are accumulated in two
    // interleaved chains per component: a third of the FMAs of the plain k = 0..18 sum and short dependency chains
    // (the kernel runs 8 warps per SM, so FP64 latency has to be hidden inside the thread)
    double gx[2] = {0.0, 0.0}, gy[2] = {0.0, 0.0}, gz[2] = {0.0, 0.0};
    int nx_ = 0, ny_ = 0, nz_ = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int ko = L19f::opp(k);
        const double *Ra = L19f::cx(k) < 0 ? Rm : (L19f::cx(k) > 0 ? Rp : R0);
        const double *Rb = L19f::cx(ko) < 0 ? Rm : (L19f::cx(ko) > 0 ? Rp : R0);
        double va = Ra[q + L19f::cy(k) * ZR + L19f::cz(k)];
        double vb = Rb[q + L19f::cy(ko) * ZR + L19f::cz(ko)];
        if (WALLS && (wall & (1u << k))) va = xc;
        if (WALLS && (wall & (1u << ko))) vb = xc;
        const double d = L19f::t(k) * (va - vb);
        if (L19f::cx(k)) { gx[nx_ & 1] += L19f::cx(k) * d; ++nx_; }
        if (L19f::cy(k)) { gy[ny_ & 1] += L19f::cy(k) * d; ++ny_; }
        if (L19f::cz(k)) { gz[nz_ & 1] += L19f::cz(k) * d; ++nz_; }
    }
    g[0] = 3.0 * (gx[0] + gx[1]);
    g[1] = 3.0 * (gy[0] + gy[1]);
    g[2] = 3.0 * (gz[0] + gz[1]);
}

// bit k set: the k-th neighbour of (a, b) (halo-3 ring coordinates) is a bounce_back node
template <int ZR>
CLBM_D unsigned wall19(const uint8_t *Fm, const uint8_t *F0, const uint8_t *Fp, int q)
{
    unsigned wall = 0;
#pragma unroll
    for (int k = 0; k < 19; ++k) {
        if (k == L19f::REST) continue;
        const uint8_t *F = L19f::cx(k) < 0 ? Fm : (L19f::cx(k) > 0 ? Fp : F0);
        if (F[q + L19f::cy(k) * ZR + L19f::cz(k)] == CELL_BB) wall |= 1u << k;
    }
    return wall;
}

// what the collision of a node needs from level 2 (kept in registers by the owning thread)
struct Hcz3dLocal {
    double phi, rho, Fx, Fy, Fz, gp[3], u0, u1, u2, Pt;
};

template <int TY, int TZ>
__global__ void __launch_bounds__(TY *TZ, (TY * TZ <= 128) ? 2 : 1)
hcz3d_fused_kernel(const __grid_constant__ CUtensorMap tmap_f, const __grid_constant__ CUtensorMap tmap_g, const Hcz3dOut P,
                   const Hcz3dMom M, const uint8_t *__restrict__ flag, Geom g, ModelParams mp, int xchunk)
{
    using C = Hcz3dCfg<TY, TZ>;
    constexpr int NT = C::NT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t stage_a = smem_u32(smem_raw);
    double *r_phi = reinterpret_cast<double *>(smem_raw + C::OFF_PHI);   // [4][Y3][Z3]
    double *r_lap = reinterpret_cast<double *>(smem_raw + C::OFF_LAP);   // [4][Y2][Z2]
    double *r_pp = reinterpret_cast<double *>(smem_raw + C::OFF_PP);     // [4][Y2][Z2]
    double *r_pr = reinterpret_cast<double *>(smem_raw + C::OFF_PR);     // [4][Y1][Z1]
    uint8_t *r_fl = smem_raw + C::OFF_FL;                                // [8][Y3][Z3]
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + C::OFF_BAR);

    const int tid = threadIdx.x;
    const int tz = tid % TZ, ty = tid / TZ;
    const int y0 = blockIdx.y * TY, z0 = blockIdx.x * TZ;
    const int y = y0 + ty, z = z0 + tz;
    const bool inside = (y < g.ny) && (z < g.nz);
    const int xa = blockIdx.z * xchunk;
    const int xb = min(g.nx, xa + xchunk);
    const int plane = (int)g.plane, nz = g.nz, ny = g.ny, G = g.G;
    const int yz = y * nz + z;
    auto wrap = [](int v, int n) { v %= n; return v < 0 ? v + n : v; };
    const int yz_w = wrap(y, ny) * nz + wrap(z, nz);   // periodic image of the own cell (ragged edge tiles)
    auto xs_of = [&](int xg) { return g.wx(xg) + G; };   // storage plane of slab plane xg

    // S1 cells of this thread (halo-3 region, row-major), as in-plane lattice offsets
    int c3_yz[C::N3];
#pragma unroll
    for (int j = 0; j < C::N3; ++j) {
        const int h = tid + j * NT;
        const int sy = h / C::Z3, sz = h % C::Z3;
        c3_yz[j] = (h < C::R3) ? wrap(y0 + sy - 3, ny) * nz + wrap(z0 + sz - 3, nz) : -1;
    }
    // halo-1 ring cell of this thread (S3), in halo-1 coordinates
    const bool h_act = tid < C::NH1;
    int h1y = 0, h1z = 0;
    if (h_act) {
        if (tid < C::Z1) { h1y = 0; h1z = tid; }
        else if (tid < 2 * C::Z1) { h1y = C::Y1 - 1; h1z = tid - C::Z1; }
        else { const int q = tid - 2 * C::Z1; h1y = 1 + (q >> 1); h1z = (q & 1) ? C::Z1 - 1 : 0; }
    }
    const bool h_warp = (tid >> 5) <= ((C::NH1 - 1) >> 5);
    if (h_warp && !h_act) { h1y = ty + 1; h1z = tz + 1; }
    const int h1_yz = wrap(y0 + h1y - 1, ny) * nz + wrap(z0 + h1z - 1, nz);

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int x) {   // populations of plane x (both sets) into stage (x - xa) & 1
        const int s = (x - xa) & 1;
        mbar_expect_tx(&mbar[s], (uint32_t)C::STAGE_BYTES);
        tma_load_4d(stage_a + s * C::STAGE_BYTES, &tmap_f, &mbar[s], z0, y0, x + G, 0);
        tma_load_4d(stage_a + s * C::STAGE_BYTES + C::SET_BYTES, &tmap_g, &mbar[s], z0, y0, x + G, 0);
    };
    if (tid == 0) {
        issue(xa);
        if (xa + 1 < xb) issue(xa + 1);
    }

    // registers that run one plane ahead of their use
    double phi_n[C::N3];
    uint8_t fl_n[C::N3];
    auto load_phi = [&](int xg) {
        const int base = xs_of(xg) * plane;
#pragma unroll
        for (int j = 0; j < C::N3; ++j)
            if (c3_yz[j] >= 0) { phi_n[j] = M.phi[base + c3_yz[j]]; fl_n[j] = flag[base + c3_yz[j]]; }
    };
    double mo_n[4] = {0., 0., 0., 0.}, mh_n[4] = {0., 0., 0., 0.}, mo_c[4], mh_c[4];
    auto load_mom = [&](int xg) {
        const int base = xs_of(xg) * plane;
        mo_n[0] = M.pt[base + yz_w]; mo_n[1] = M.jx[base + yz_w]; mo_n[2] = M.jy[base + yz_w]; mo_n[3] = M.jz[base + yz_w];
        if (h_warp) { mh_n[0] = M.pt[base + h1_yz]; mh_n[1] = M.jx[base + h1_yz]; mh_n[2] = M.jy[base + h1_yz]; mh_n[3] = M.jz[base + h1_yz]; }
    };

    // level 2 of the node at halo-1 position (a1, b1) of plane p; stores psi_rho, returns the node's local set
    auto level2 = [&](auto wtag, int p, int a1, int b1, const double *mo, Hcz3dLocal &o) -> double {
        constexpr bool W = decltype(wtag)::value;
        const int q3 = (a1 + 2) * C::Z3 + (b1 + 2), q2 = (a1 + 1) * C::Z2 + (b1 + 1);
        const int sm = ((p - 1) & 3) * C::R2, s0 = (p & 3) * C::R2, sp = ((p + 1) & 3) * C::R2;
        double gl[3];
        unsigned wall = 0;
        if constexpr (W) {
            const uint8_t *Fm = r_fl + ((p - 1) & 7) * C::R3, *F0 = r_fl + (p & 7) * C::R3, *Fp = r_fl + ((p + 1) & 7) * C::R3;
            wall = wall19<C::Z3>(Fm, F0, Fp, q3);
        }
        grad19<C::Z2, W>(r_lap + sm, r_lap + s0, r_lap + sp, q2, wall, gl);
        grad19<C::Z2, W>(r_pp + sm, r_pp + s0, r_pp + sp, q2, wall, o.gp);
        o.phi = r_phi[(p & 3) * C::R3 + q3];
        o.rho = mp.rho_g + ((o.phi - mp.phi_g) * mp.inv_dphi) * mp.drho;
        o.Fx = mp.kappa * o.phi * gl[0];
        o.Fy = mp.kappa * o.phi * gl[1] + mp.gravity * o.rho;
        o.Fz = mp.kappa * o.phi * gl[2];
        const double inv_d = 3.0 * fast_rcp(o.rho);   // 1 / (rho/3)
        o.u0 = (mo[1] + o.Fx * (1. / 6.)) * inv_d;
        o.u1 = (mo[2] + o.Fy * (1. / 6.)) * inv_d;
        o.u2 = (mo[3] + o.Fy * (1. / 6.)) * inv_d;   // sic: forcey (laplace3D.h:304)
        o.Pt = mo[0] - 0.5 * (o.u0 * o.gp[0] + o.u1 * o.gp[1] + o.u2 * o.gp[2]);
        return o.Pt - o.rho * (1. / 3.);   // psi_rho; the caller stores it (after all its evaluations: the loads of one
                                           // evaluation must not wait behind the store of the previous one)
    };

    const int oym = (g.wy(y - 1) - y) * nz, oyp = (g.wy(y + 1) - y) * nz;
    const int ozm = g.wz(z - 1) - z, ozp = g.wz(z + 1) - z;
    const double omega = mp.omega, om1 = 1. - omega, hw = 1. - 0.5 * omega;

    Hcz3dLocal cur, nxt;
    cur.phi = cur.rho = 1.0; cur.Fx = cur.Fy = cur.Fz = 0.0; cur.gp[0] = cur.gp[1] = cur.gp[2] = 0.0;
    cur.u0 = cur.u1 = cur.u2 = cur.Pt = 0.0;
    nxt = cur;

    // S2 cells of this thread as ring offsets (halo-3 / halo-2 coordinates)
    // (a lane past the region in a partly active warp repeats its first cell, so that every round is warp-uniform)
    int c2_q3[C::N2], c2_q2[C::N2];
    bool c2_on[C::N2];
#pragma unroll
    for (int j = 0; j < C::N2; ++j) {
        int h = tid + j * NT;
        c2_on[j] = ((tid & ~31) + j * NT) < C::R2;   // warp-uniform
        if (h >= C::R2) h = tid;
        const int a2 = h / C::Z2, b2 = h % C::Z2;
        c2_q2[j] = h;
        c2_q3[j] = (a2 + 1) * C::Z3 + (b2 + 1);
    }
    unsigned wmask = 0;   // bit (p & 7): plane p has a bounce_back node inside this CTA's window (CTA-uniform)

    // S4 of the own node: grad psi_rho, the 2 x 19 collisions, push.  Two instantiations: with walls in reach the
    // destination of every population is selected without a branch (bounce-back: opposite slot of the own node), the
    // wall-free one has no mask work at all; straight-line code in both, so the scheduler can overlap the FP64
    // chains of neighbouring directions (8 warps per SM: latency has to be hidden inside the thread)
    auto collide = [&](auto wtag, int x, const double *sf, const uint8_t *Fm, const uint8_t *F0, const uint8_t *Fp) {
        constexpr bool W = decltype(wtag)::value;
        constexpr int R1 = C::Y1 * C::Z1;
        const int q1 = (ty + 1) * C::Z1 + tz + 1;
        unsigned wall = 0;
        double ge[3];
        if constexpr (W) wall = wall19<C::Z3>(Fm, F0, Fp, (ty + 3) * C::Z3 + tz + 3);
        grad19<C::Z1, W>(r_pr + ((x - 1) & 3) * R1, r_pr + (x & 3) * R1, r_pr + ((x + 1) & 3) * R1, q1, wall, ge);

        const double phi = cur.phi, rho = cur.rho;
        const double u0 = cur.u0, u1 = cur.u1, u2 = cur.u2, Pt = cur.Pt;
        const double usqr = 1.5 * (u0 * u0 + u1 * u1 + u2 * u2);
        // The two forcing terms, regrouped around Gamma_k = eqf_k / phi = t_k (1 + poly_k):
        //   fg_k = hw [ (c_k - u).F Gamma_k + (c_k - u).(-E) (Gamma_k - t_k) ] = Gamma_k (c_k - u).D + t_k (c_k - u).Eh
        //   ff_k = hw (c_k - u).(-grad psi(phi)) 3 eqf_k / rho                 = Gamma_k (c_k - u).Gv
        // with the per-node vectors D = hw (F - E), Eh = hw E, Gv = -(3 hw phi / rho) grad psi(phi)
        const double ffs = -hw * 3.0 * phi * fast_rcp(rho);
        const double D0 = hw * (cur.Fx - ge[0]), D1 = hw * (cur.Fy - ge[1]), D2 = hw * (cur.Fz - ge[2]);
        const double E0 = hw * ge[0], E1 = hw * ge[1], E2 = hw * ge[2];
        const double G0 = ffs * cur.gp[0], G1 = ffs * cur.gp[1], G2 = ffs * cur.gp[2];
        const double uD = u0 * D0 + u1 * D1 + u2 * D2;
        const double uE = u0 * E0 + u1 * E1 + u2 * E2;
        const double opg = omega * phi - (u0 * G0 + u1 * G1 + u2 * G2);   // omega phi - u.Gv
        const double rho3 = rho * (1. / 3.);
        // omega t_k (Pt + rho/3 poly_k) = A + B poly_k with one (A, B) pair per weight class
        const double Aa = (omega * (1. / 18.)) * Pt, Ba = (omega * (1. / 18.)) * rho3;
        const double Ad = (omega * (1. / 36.)) * Pt, Bd = (omega * (1. / 36.)) * rho3;
        const int xp = g.wx(x + 1), xm = g.wx(x - 1);
        const int i = (x + G) * plane + yz;
        const int oxm = (xm - x) * plane, oxp = (xp - x) * plane;
#pragma unroll
        for (int k = 0; k < 19; ++k) {
            const double fk = sf[k * NT];
            const double gk = sf[(19 + k) * NT];
            const double t = L19f::t(k);
            double pf, pg;
            if (k == 9) {
                const double Gam = fma(-t, usqr, t);                  // eqf0 / phi = t (1 - usqr)
                pf = fma(Gam, opg, om1 * fk);
                pg = fma(om1, gk, (omega * t) * fma(-rho3, usqr, Pt)) - fma(Gam, uD, t * uE);
                P.fout[k][i] = pf;
                P.gout[k][i] = pg;
                continue;
            }
            const bool axis = (L19f::cx(k) != 0) + (L19f::cy(k) != 0) + (L19f::cz(k) != 0) == 1;
            const double cu = cdot<L19f>(k, u0, u1, u2);
            const double poly = fma(cu, fma(4.5, cu, 3.0), -usqr);    // 3 cu + 4.5 cu^2 - usqr
            const double Gam = fma(t, poly, t);                       // eqf / phi = t (1 + poly)
            const double dD = cdot<L19f>(k, D0, D1, D2) - uD;
            const double dE = cdot<L19f>(k, E0, E1, E2) - uE;
            const double dG = cdot<L19f>(k, G0, G1, G2) + opg;
            pf = fma(Gam, dG, om1 * fk);
            pg = fma(t, dE, fma(Gam, dD, fma(om1, gk, fma(axis ? Ba : Bd, poly, axis ? Aa : Ad))));
            const int off = (L19f::cx(k) < 0 ? oxm : (L19f::cx(k) > 0 ? oxp : 0)) + (L19f::cy(k) < 0 ? oym : (L19f::cy(k) > 0 ? oyp : 0)) +
                            (L19f::cz(k) < 0 ? ozm : (L19f::cz(k) > 0 ? ozp : 0));
            if constexpr (W) {
                const bool bb = (wall >> k) & 1u;
                double *df = bb ? P.fout[L19f::opp(k)] + i : P.fout[k] + (i + off);
                double *dg = bb ? P.gout[L19f::opp(k)] + i : P.gout[k] + (i + off);
                *df = pf;
                *dg = pg;
            } else {
                P.fout[k][i + off] = pf;
                P.gout[k][i + off] = pg;
            }
        }
    };

    load_phi(xa - 3);
    // x = plane being collided; the first six iterations only fill the pipeline (x < xa)
    for (int x = xa - 6; x < xb; ++x) {
        // ---- S1: phi / mask of plane x+3 from the registers, then prefetch the next plane's ----
        int anyw = 0;
        {
            double *dst = r_phi + ((x + 3) & 3) * C::R3;
            uint8_t *dfl = r_fl + ((x + 3) & 7) * C::R3;
#pragma unroll
            for (int j = 0; j < C::N3; ++j)
                if (c3_yz[j] >= 0) { dst[tid + j * NT] = phi_n[j]; dfl[tid + j * NT] = fl_n[j]; anyw |= (fl_n[j] == CELL_BB); }
        }
        if (x + 1 < xb) load_phi(x + 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) { mo_c[j] = mo_n[j]; mh_c[j] = mh_n[j]; }
        if (x + 2 >= xa - 1 && x + 1 < xb) load_mom(x + 2);   // consumed by S3 of the next iteration
        {
            const unsigned bit = 1u << ((x + 3) & 7);
            wmask = __syncthreads_or(anyw) ? (wmask | bit) : (wmask & ~bit);
        }
        auto walls_near = [&](int p) { return (wmask & ((1u << ((p - 1) & 7)) | (1u << (p & 7)) | (1u << ((p + 1) & 7)))) != 0u; };
        if (tid == 0 && x - 1 >= xa && x + 1 < xb) issue(x + 1);   // stage of plane x-1 is free now

        // ---- S2: lap(phi), psi(phi) of plane x+2 on tile + halo 2 ----
        if (x + 2 >= xa - 2) {
            const int p = x + 2;
            const double *Pm = r_phi + ((p - 1) & 3) * C::R3, *P0 = r_phi + (p & 3) * C::R3, *Pp = r_phi + ((p + 1) & 3) * C::R3;
            const uint8_t *Fm = r_fl + ((p - 1) & 7) * C::R3, *F0 = r_fl + (p & 7) * C::R3, *Fp = r_fl + ((p + 1) & 7) * C::R3;
            // all cells of the thread are evaluated before the first store (the loads of one cell must not wait behind the
            // stores of the previous one), in a walls / wall-free instantiation
            auto s2 = [&](auto wtag) {
                constexpr bool W = decltype(wtag)::value;
                double lap_v[C::N2], pp_v[C::N2];
#pragma unroll
                for (int j = 0; j < C::N2; ++j) {
                    if (!c2_on[j]) continue;
                    const int q3 = c2_q3[j];
                    const double phi_c = P0[q3];
                    double sum = 0.0;
                    if constexpr (W) {
#pragma unroll
                        for (int k = 0; k < 19; ++k) {
                            if (k == L19f::REST) continue;
                            const double *R = L19f::cx(k) < 0 ? Pm : (L19f::cx(k) > 0 ? Pp : P0);
                            const uint8_t *F = L19f::cx(k) < 0 ? Fm : (L19f::cx(k) > 0 ? Fp : F0);
                            const int q = q3 + L19f::cy(k) * C::Z3 + L19f::cz(k);
                            if (F[q] != CELL_BB) sum += L19f::t(k) * (R[q] - phi_c);
                        }
                    } else {
                        // no wall in reach: sum_k t_k (phi_nb - phi_c) = (1/18) S_axis + (1/36) S_diag - (2/3) phi_c with the
                        // neighbour values added in four independent chains
                        double sa[2] = {0.0, 0.0}, sd[2] = {0.0, 0.0};
                        int na = 0, nd = 0;
#pragma unroll
                        for (int k = 0; k < 19; ++k) {
                            if (k == L19f::REST) continue;
                            const double *R = L19f::cx(k) < 0 ? Pm : (L19f::cx(k) > 0 ? Pp : P0);
                            const double v = R[q3 + L19f::cy(k) * C::Z3 + L19f::cz(k)];
                            const bool axis = (L19f::cx(k) != 0) + (L19f::cy(k) != 0) + (L19f::cz(k) != 0) == 1;
                            if (axis) { sa[na & 1] += v; ++na; }
                            else { sd[nd & 1] += v; ++nd; }
                        }
                        sum = (1. / 18.) * (sa[0] + sa[1]) + (1. / 36.) * (sd[0] + sd[1]) - (2. / 3.) * phi_c;
                    }
                    lap_v[j] = 6.0 * sum;
                    pp_v[j] = hcz_psi1(phi_c, mp.a, mp.b);
                }
#pragma unroll
                for (int j = 0; j < C::N2; ++j) {
                    if (!c2_on[j]) continue;
                    r_lap[(p & 3) * C::R2 + c2_q2[j]] = lap_v[j];
                    r_pp[(p & 3) * C::R2 + c2_q2[j]] = pp_v[j];
                }
            };
            if (walls_near(p)) s2(std::true_type{}); else s2(std::false_type{});
        }
        __syncthreads();

        // ---- S3: level 2 of plane x+1 on tile + halo 1 ----
        if (x + 1 >= xa - 1) {
            const bool walls = walls_near(x + 1);
            // the halo-1 ring belongs to the first warps; they evaluate both nodes in one basic block so that the two
            // independent dependency chains overlap (lanes past the ring repeat their own node, the stores coincide)
            auto s3 = [&](auto wtag) {
                double *pr = r_pr + ((x + 1) & 3) * (C::Y1 * C::Z1);
                if (h_warp) {
                    Hcz3dLocal tmp;
                    const double a = level2(wtag, x + 1, ty + 1, tz + 1, mo_c, nxt);
                    const double b = level2(wtag, x + 1, h1y, h1z, mh_c, tmp);
                    pr[(ty + 1) * C::Z1 + tz + 1] = a;
                    pr[h1y * C::Z1 + h1z] = b;
                } else {
                    pr[(ty + 1) * C::Z1 + tz + 1] = level2(wtag, x + 1, ty + 1, tz + 1, mo_c, nxt);   // also the periodic images in a ragged edge tile
                }
            };
            if (walls) s3(std::true_type{}); else s3(std::false_type{});
        }
        __syncthreads();

        // ---- S4: collide + push plane x ----
        if (x >= xa) {
            const int r = x - xa;
            mbar_wait(&mbar[r & 1], (r >> 1) & 1);
            const uint8_t *Fm = r_fl + ((x - 1) & 7) * C::R3, *F0 = r_fl + (x & 7) * C::R3, *Fp = r_fl + ((x + 1) & 7) * C::R3;
            if (inside && F0[(ty + 3) * C::Z3 + tz + 3] == CELL_BULK) {
                const double *sf = reinterpret_cast<const double *>(smem_raw + (r & 1) * C::STAGE_BYTES) + tid;
                if (walls_near(x)) collide(std::true_type{}, x, sf, Fm, F0, Fp);
                else collide(std::false_type{}, x, sf, Fm, F0, Fp);
            }
        }
        cur = nxt;
    }
}

bool hcz3d_fused_eligible(const clbm_ctx *c)
{
    const Geom &g = c->geo;
    return (g.nz % 2 == 0) && g.nx >= 4 && g.ny >= 4 && g.nz >= 4 && g.ncs < (1LL << 31) && get_encode() != nullptr;
}

template <int TY, int TZ>
static int launch_hcz3d_fused(clbm_ctx *c)
{
    using C = Hcz3dCfg<TY, TZ>;
    const Geom &g = c->geo;
    CUtensorMap tm[2];
    const cuuint32_t box[4] = {(cuuint32_t)TZ, (cuuint32_t)TY, 1, 19};
    for (int s = 0; s < 2; ++s)
        if (int rc = cached_tmap(c, c->pop[s][c->parity], box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, &tm[s])) return rc;
    const int tiles = ((g.ny + TY - 1) / TY) * ((g.nz + TZ - 1) / TZ);
    // every chunk pays a 6-plane pipeline fill, so chunks stay long here (128 planes measured best at 512^3)
    int xchunk = g.nx < 128 ? g.nx : 128;
    const long long want = 4LL * 148;
    if ((long long)tiles * ((g.nx + xchunk - 1) / xchunk) < want) {
        const long long nch = (want + tiles - 1) / tiles;
        xchunk = (int)((g.nx + nch - 1) / nch);
        if (xchunk < 32) xchunk = g.nx < 32 ? g.nx : 32;
    }
    if (c->env.hcz_xchunk > 0) xchunk = c->env.hcz_xchunk < g.nx ? c->env.hcz_xchunk : g.nx;
    dim3 grid((g.nz + TZ - 1) / TZ, (g.ny + TY - 1) / TY, (g.nx + xchunk - 1) / xchunk);
    Hcz3dOut P;
    for (int k = 0; k < 19; ++k) {
        P.fout[k] = c->pop[0][1 - c->parity] + (size_t)k * g.ncs;
        P.gout[k] = c->pop[1][1 - c->parity] + (size_t)k * g.ncs;
    }
    const Hcz3dMom M = {c->fld[0], c->fld[1], c->fld[2], c->fld[3], c->fld[4]};
    auto kern = hcz3d_fused_kernel<TY, TZ>;
    static PerDeviceOnce attr;
    if (attr.need(c->device)) {
        CLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr.mark(c->device);
    }
    LaunchScope ls(c, "hcz3d_fused_collide_stream", true);
    kern<<<grid, C::NT, C::SMEM, c->stream>>>(tm[0], tm[1], P, M, c->flag, g, c->mp, xchunk);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

int hcz3d_fused_launch(clbm_ctx *c, int variant)
{
    switch (variant) {
    case 8: return launch_hcz3d_fused<16, 16>(c);
    case 9: return launch_hcz3d_fused<4, 64>(c);
    case 10: return launch_hcz3d_fused<8, 16>(c);   // two 128-thread CTAs per SM
    default: return launch_hcz3d_fused<8, 32>(c);
    }
}

}  // namespace clbm
