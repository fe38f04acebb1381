// hcz3d_sweep.cu -- HCZ D3Q19 (PF/apps/laplace3D.h): the WHOLE time step in ONE plane-marching sweep (sm_100a).
//
// hcz3d_fused.cu needs a separate pass over all 38 populations (hcz3d_moments_kernel) before it can start, because the
// stencil chain phi -> lap phi -> grad lap phi -> u -> total_P -> psi_rho -> grad psi_rho (:216-336, :370-500) reaches
// three nodes and the reference pushes: the moments of a node are sums over populations that ARRIVED.  That pass reads 304 B
// per node on top of the 609 algorithmic bytes.  Here the moments of the NEXT step are accumulated while this step pushes:
//
//   S1  phi of plane x+3 on the tile + halo 3 (moment array of the previous sweep + edge sums, hcz3d_edges.cuh)
//   S2  lap(phi) (:370-393) and psi(phi) (:268-275) of plane x+2 on tile + halo 2
//   S3  plane x+1 on tile + halo 1: grad lap(phi), grad psi(phi) (:435-536), velocity with the forcey-in-z quirk (:280-312),
//       total_P (:318-328), psi_rho (:330-336); the owning thread keeps the node's local set in registers
//   S4  plane x: grad psi_rho, collideBgk of the 2 x 19 populations (:562-624, rest :664-677), push (:539-559); the
//       post-collision values are also written back IN PLACE into the TMA stage the inputs came from
//   S5  (one iteration later, next to S2) every node -- and every cell of the one-cell ring around the tile -- gathers what
//       the tile's nodes of plane x pushed into it, grouped by the x component of the direction:
//       A (c_x = +1, lands in plane x+1), B (c_x = 0, plane x), C (c_x = -1, plane x-1).  Plane x-1 is then complete:
//       (A + B) + C, with the two earlier groups carried in thread-private shared-memory slots, and goes to the moment
//       arrays of the next step (tile nodes) or to the edge arrays (ring cells: single writer per slot, plain stores).
//
// Wall-free lattices only (the reference's laplace3D has no walls: inigeom :830-849 never fires); a lattice with bounce_back
// nodes, a ragged tile grid or too few tiles runs the two-pass path of hcz3d_fused.cu.  One CTA marches the whole x range
// of its tile (no x-chunks), so the only partial sums are the y/z ring; x wraps inside the CTA (the first plane's C group
// and the last plane's A group are merged into planes nx-1 and 0 at the end of the march).
#include <cstdlib>

#include "hcz3d_edges.cuh"
#include "sc_cell.cuh"
#include "tma.cuh"

namespace clbm {

using L19s = D3Q19;

// out buffers of the two population sets: direction k starts at k * ncs.  (An array of 38 pointers in the parameter block costs
// a constant-bank load per store -- they cannot all stay in registers -- and the store then waits on that load: ncu showed
// the long-scoreboard stalls of this kernel sitting on exactly those address computations.)
struct SweepOut {
    double *f, *g;
    unsigned kn[19];   // k * ncs as unsigned 32-bit element offsets (hcz3d_sweep_shape_ok: 19 * ncs < 2^32): the 38 stores of a node
                       // cost two or three integer instructions each instead of seven for k * ncs + i + offset in 64 bits
};
// phi, P_term, jx, jy, jz: node arrays m[5] of [ncs]; edge sums of phi in ephi[nx][eplane]; edge sums of the other four
// moments INTERLEAVED in e4[nx][eplane][4] (one address, four consecutive loads / two 16-byte stores per slot)
struct SweepMom { double *m[5]; double *ephi; double *e4; };

template <int TY, int TZ>
struct SweepCfg {
    static constexpr int NT = TY * TZ;
    static constexpr int Y3 = TY + 6, Z3 = TZ + 6, R3 = Y3 * Z3;   // tile + halo 3
    static constexpr int Y2 = TY + 4, Z2 = TZ + 4, R2 = Y2 * Z2;   // tile + halo 2
    static constexpr int Y1 = TY + 2, Z1 = TZ + 2, R1 = Y1 * Z1;   // tile + halo 1
    static constexpr int NH1 = 2 * Z1 + 2 * TY;                    // cells of the halo-1 ring
    static constexpr int N2 = (R2 + NT - 1) / NT;                  // S2 cells per thread
    static constexpr int NSPEC = 4 * Z3 + 4 * (Y3 - 4);            // halo-3 window cells that are border nodes of their tile
    static constexpr int NPLAIN = R3 - NSPEC;
    static constexpr int NACC = NT + NH1;                          // accumulator slots: tile nodes, then ring cells
    static constexpr int SET_BYTES = 19 * NT * 8;
    static constexpr int STAGE_BYTES = 2 * SET_BYTES;
    static constexpr int OFF_PHI = 2 * STAGE_BYTES;                // [4][R3]
    static constexpr int OFF_LAP = OFF_PHI + 4 * R3 * 8;           // [3][R2]
    static constexpr int OFF_PP = OFF_LAP + 3 * R2 * 8;            // [3][R2]
    static constexpr int OFF_PR = OFF_PP + 3 * R2 * 8;             // [4][R1]
    static constexpr int OFF_ACC = OFF_PR + 4 * R1 * 8;            // [9][NACC]
    static constexpr int OFF_BAR = ((OFF_ACC + 9 * NACC * 8 + 15) / 16) * 16;
    static constexpr int SMEM = OFF_BAR + 32;
    static_assert(NH1 <= NT, "one ring cell per thread");
    static_assert(NSPEC <= NT && NPLAIN <= 2 * NT, "window cells per thread");
    static_assert(TZ % 2 == 0 && SET_BYTES % 128 == 0, "TMA alignment");
    static_assert((TY & (TY - 1)) == 0 && (TZ & (TZ - 1)) == 0 && TY >= 4 && TZ >= 4, "tile extents are powers of two");
    static_assert(SMEM <= 232448, "shared memory budget");
};

// 3 * sum_k t_k c_k X(nb) without walls: opposite directions paired, two interleaved chains per component
template <int ZR>
CLBM_D void grad19s(const double *Rm, const double *R0, const double *Rp, int q, double g[3])
{
    double gx[2] = {0.0, 0.0}, gy[2] = {0.0, 0.0}, gz[2] = {0.0, 0.0};
    int nx_ = 0, ny_ = 0, nz_ = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int ko = L19s::opp(k);
        const double *Ra = L19s::cx(k) < 0 ? Rm : (L19s::cx(k) > 0 ? Rp : R0);
        const double *Rb = L19s::cx(ko) < 0 ? Rm : (L19s::cx(ko) > 0 ? Rp : R0);
        const double va = Ra[q + L19s::cy(k) * ZR + L19s::cz(k)];
        const double vb = Rb[q + L19s::cy(ko) * ZR + L19s::cz(ko)];
        const double d = L19s::t(k) * (va - vb);
        if (L19s::cx(k)) { gx[nx_ & 1] += L19s::cx(k) * d; ++nx_; }
        if (L19s::cy(k)) { gy[ny_ & 1] += L19s::cy(k) * d; ++ny_; }
        if (L19s::cz(k)) { gz[nz_ & 1] += L19s::cz(k) * d; ++nz_; }
    }
    g[0] = 3.0 * (gx[0] + gx[1]);
    g[1] = 3.0 * (gy[0] + gy[1]);
    g[2] = 3.0 * (gz[0] + gz[1]);
}

struct SweepLocal {
    double phi, rho, Fx, Fy, Fz, gp[3], u0, u1, u2, Pt;
};

// VAR (CLBM_HCZ3D_SWEEP_VAR, bit-identical results): bit 0 = the four interleaved edge moments of a slot as two 16-byte loads
// instead of four 8-byte ones (the ring warps issue 38 scattered loads per plane in S1 and sat in the LSU queue there: `lg`
// throttle on exactly those lines); bit 1 = the raw moments (needed after the gather phase) are requested BEFORE phi of plane
// x+4 (needed an iteration later); bit 2 = the push of the nine directions with c_z = 0 as TMA box stores: S4 writes the post-collision values back
// into the stage anyway (S5 gathers them there), so one thread stores the tile of such a direction to its shifted position with
// cp.async.bulk.tensor shared -> global boxes (18 per plane) instead of a st.global per thread and direction (the knock-out timing
// of tools/hcz3d_sweep_ko.py: 4.8 of 22.3 ms per step went with the thread-level stores).  Only c_z = 0: a box STORE traps with
// "illegal instruction" on this device unless its start is 16-byte aligned in global memory and its coordinates are non-negative
// (tools/probe/tma_store_probe.cu; box LOADS have neither restriction), and z0 +- 1 is an odd number of doubles.  Tiles in the
// first / last tile row (y0 - 1 < 0, or a clipped box) keep the thread-level stores for every direction.
template <int TY, int TZ, int VAR = 0, bool KO = false>
__global__ void __launch_bounds__(TY *TZ, 1)
hcz3d_sweep_kernel(const __grid_constant__ CUtensorMap tmap_f, const __grid_constant__ CUtensorMap tmap_g,
                   const __grid_constant__ CUtensorMap tmap_fo, const __grid_constant__ CUtensorMap tmap_go, const SweepOut P,
                   const SweepMom Min, const SweepMom Mout, Geom g, ModelParams mp, EdgeGeom eg, int ko_arg)
{
    // ko: knock-out bits of tools/hcz3d_sweep_ko.py (CLBM_HCZ3D_SWEEP_KO) -- TIMING EXPERIMENTS ONLY, results are wrong when set; only the
    // KO = true instantiation looks at them, in the production kernels the constant 0 removes every test
    const int ko = KO ? ko_arg : 0;
    using C = SweepCfg<TY, TZ>;
    constexpr int NT = C::NT;
    constexpr bool TS = (VAR & 4) != 0;             // TMA stores
    constexpr int TMA_TID = TS ? NT - 32 : 0;       // the thread that talks to the TMA unit (TS: first lane of the last warp, which has no ring cells)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t stage_a = smem_u32(smem_raw);
    double *r_phi = reinterpret_cast<double *>(smem_raw + C::OFF_PHI);   // [4][R3]
    double *r_lap = reinterpret_cast<double *>(smem_raw + C::OFF_LAP);   // [3][R2]
    double *r_pp = reinterpret_cast<double *>(smem_raw + C::OFF_PP);     // [3][R2]
    double *r_pr = reinterpret_cast<double *>(smem_raw + C::OFF_PR);     // [4][R1]
    double *acc = reinterpret_cast<double *>(smem_raw + C::OFF_ACC);     // [9][NACC]: T (5) then A (4), thread-private slots
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + C::OFF_BAR);

    const int tid = threadIdx.x;
    const int tz = tid % TZ, ty = tid / TZ;
    const int y0 = blockIdx.y * TY, z0 = blockIdx.x * TZ;
    const int y = y0 + ty, z = z0 + tz;
    const int nx = g.nx, plane = (int)g.plane, nz = g.nz, ny = g.ny, G = g.G;
    const bool wrapx = g.wrapx != 0;
    const int yz = y * nz + z;
    auto wrapn = [](int v, int n) { return v < 0 ? v + n : (v >= n ? v - n : v); };
    auto mod3 = [](int p) { return (p + 30) % 3; };

    // ---- halo-3 window cells of this thread: slots 0, 1 = nodes that are NOT on a tile border, slot 2 = border nodes,
    //      whose phi is the node array + up to three edge sums ----
    int w_idx[3], w_yz[3], w_e[3] = {-1, -1, -1};
    {
        auto plain_row = [](int ri) { return ri < 2 ? ri : (ri < TY ? ri + 2 : ri + 4); };
        auto plain_col = [](int ci) { return ci < 2 ? ci : (ci < TZ ? ci + 2 : ci + 4); };
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int n = tid + j * NT;
            w_idx[j] = -1;
            w_yz[j] = 0;
            if (n < C::NPLAIN) {
                const int sy = plain_row(n / (TZ + 2)), sz = plain_col(n % (TZ + 2));
                w_idx[j] = sy * C::Z3 + sz;
                w_yz[j] = wrapn(y0 + sy - 3, ny) * nz + wrapn(z0 + sz - 3, nz);
            }
        }
        w_idx[2] = -1;
        w_yz[2] = 0;
        if (tid < C::NSPEC) {
            int sy, sz;
            if (tid < 4 * C::Z3) {
                const int r = tid / C::Z3;
                sy = (r < 2 ? 2 + r : TY + r);          // rows 2, 3, TY+2, TY+3
                sz = tid % C::Z3;
            } else {
                const int t = tid - 4 * C::Z3, cc = t / (TY + 2);
                sz = (cc < 2 ? 2 + cc : TZ + cc);       // columns 2, 3, TZ+2, TZ+3
                sy = plain_row(t % (TY + 2));
            }
            const int yy = wrapn(y0 + sy - 3, ny), zz = wrapn(z0 + sz - 3, nz);
            w_idx[2] = sy * C::Z3 + sz;
            w_yz[2] = yy * nz + zz;
            edge_offsets<TY, TZ>(eg, yy, zz, w_e);
        }
    }
    // ---- the ring cell of this thread (first NH1 threads), in halo-1 coordinates; its edge slot; the edge slots of the
    //      P_term / momentum reads of the own node and of the ring cell ----
    const bool h_act = tid < C::NH1;
    int h1y = 0, h1z = 0;
    if (h_act) {
        if (tid < C::Z1) { h1y = 0; h1z = tid; }
        else if (tid < 2 * C::Z1) { h1y = C::Y1 - 1; h1z = tid - C::Z1; }
        else { const int q = tid - 2 * C::Z1; h1y = 1 + (q >> 1); h1z = (q & 1) ? C::Z1 - 1 : 0; }
    }
    const bool h_warp = (tid >> 5) <= ((C::NH1 - 1) >> 5);
    if (h_warp && !h_act) { h1y = ty + 1; h1z = tz + 1; }
    const int h1_yy = wrapn(y0 + h1y - 1, ny), h1_zz = wrapn(z0 + h1z - 1, nz);
    const int h1_yz = h1_yy * nz + h1_zz;
    int own_e[3], h1_e[3];
    edge_offsets<TY, TZ>(eg, y, z, own_e);
    edge_offsets<TY, TZ>(eg, h1_yy, h1_zz, h1_e);
    const int ring_e = h_act ? ring_slot<TY, TZ>(eg, y0, z0, h1y - 1, h1z - 1) : -1;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // accumulators start from zero (group A of the plane before the first one does not exist yet)
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        acc[j * C::NACC + tid] = 0.0;
        if (h_act) acc[j * C::NACC + NT + tid] = 0.0;
    }
    __syncthreads();
    auto issue = [&](int x) {   // populations of plane x (both sets) into stage x & 1
        const int s = x & 1;
        mbar_expect_tx(&mbar[s], (uint32_t)C::STAGE_BYTES);
        tma_load_4d(stage_a + s * C::STAGE_BYTES, &tmap_f, &mbar[s], z0, y0, x + G, 0);
        tma_load_4d(stage_a + s * C::STAGE_BYTES + C::SET_BYTES, &tmap_g, &mbar[s], z0, y0, x + G, 0);
    };
    if (tid == TMA_TID) {
        issue(0);
        if (1 < nx) issue(1);
    }
    // TS: tiles of the first / last tile row keep thread-level stores for every direction
    const bool edge_cta = TS && !push_by_box<TY, TZ>(y0, ny, 0);

    // storage plane of slab plane xg, and whether that plane has edge sums (ghost planes of an x-slab carry merged values)
    auto xs_of = [&](int xg) { return g.wx(xg) + G; };
    auto has_edges = [&](int xg) { return wrapx || (xg >= 0 && xg < nx); };

    // ---- phi: registers that run one plane ahead of their use.  The edge sums of a border node stay RAW in registers and are
    //      added when the plane is stored one iteration later: an add right behind the loads would wait for them every plane
    //      (measured: the kernel ran at half speed that way) ----
    double phi_n[3] = {0., 0., 0.}, phi_e[3] = {0., 0., 0.};
    auto load_phi = [&](int xg) {
        const int base = xs_of(xg) * plane;
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (w_idx[j] >= 0) phi_n[j] = Min.m[0][base + w_yz[j]];
        if (w_idx[2] >= 0 && has_edges(xg) && !(ko & 64)) {
            const double *E = Min.ephi + (size_t)g.wx(xg) * eg.eplane;
            phi_e[0] = phi_e[1] = phi_e[2] = 0.0;
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (w_e[q] >= 0) phi_e[q] = E[w_e[q]];      // (an if, not a ?: -- the select of a ?: waits for the load at once)
        } else {
            phi_e[0] = phi_e[1] = phi_e[2] = 0.0;
        }
    };
    // ---- P_term and raw momentum (moments 1..4) of the own node and of the ring cell: loaded RAW (node array + edge slots) at
    //      the top of the iteration that uses them and merged after the gather phase, i.e. with a phase of independent work
    //      between load and use; (an L2 prefetch one plane earlier was tried: its instructions and LSU queue stalls cost more than the latency saved).
    //      (Holding the raw values across the collide phase instead would cost 32 more registers there.) ----
    double mo_r[4][4], mh_r[4][4], mo_c[4], mh_c[4];
    auto load_mom = [&](int xg) {
        const int base = xs_of(xg) * plane;
        const bool ed = has_edges(xg) && !(ko & 64);
        const double *E = Min.e4 + (size_t)g.wx(xg) * eg.eplane * 4;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            mo_r[m][0] = Min.m[m + 1][base + yz];
            mo_r[m][1] = mo_r[m][2] = mo_r[m][3] = 0.0;
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (ed && own_e[q] >= 0) {
                const double *Eq = E + (size_t)own_e[q] * 4;
                if constexpr (VAR & 1) {
                    const double2 a = reinterpret_cast<const double2 *>(Eq)[0], b = reinterpret_cast<const double2 *>(Eq)[1];
                    mo_r[0][q + 1] = a.x; mo_r[1][q + 1] = a.y; mo_r[2][q + 1] = b.x; mo_r[3][q + 1] = b.y;
                } else {
#pragma unroll
                    for (int m = 0; m < 4; ++m) mo_r[m][q + 1] = Eq[m];
                }
            }
        if (h_warp && !(ko & 8)) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                mh_r[m][0] = Min.m[m + 1][base + h1_yz];
                mh_r[m][1] = mh_r[m][2] = mh_r[m][3] = 0.0;
            }
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (ed && h1_e[q] >= 0) {
                    const double *Eq = E + (size_t)h1_e[q] * 4;
                    if constexpr (VAR & 1) {
                        const double2 a = reinterpret_cast<const double2 *>(Eq)[0], b = reinterpret_cast<const double2 *>(Eq)[1];
                        mh_r[0][q + 1] = a.x; mh_r[1][q + 1] = a.y; mh_r[2][q + 1] = b.x; mh_r[3][q + 1] = b.y;
                    } else {
#pragma unroll
                        for (int m = 0; m < 4; ++m) mh_r[m][q + 1] = Eq[m];
                    }
                }
        }
    };
    auto merge_mom = [&]() {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            mo_c[m] = ((mo_r[m][0] + mo_r[m][1]) + mo_r[m][2]) + mo_r[m][3];
            if (h_warp) mh_c[m] = ((mh_r[m][0] + mh_r[m][1]) + mh_r[m][2]) + mh_r[m][3];
        }
    };

    // level 2 of the node at halo-1 position (a1, b1) of plane p; returns psi_rho, fills the node's local set
    auto level2 = [&](int p, int a1, int b1, const double *mo, SweepLocal &o) -> double {
        const int q3 = (a1 + 2) * C::Z3 + (b1 + 2), q2 = (a1 + 1) * C::Z2 + (b1 + 1);
        const int sm = mod3(p - 1) * C::R2, s0 = mod3(p) * C::R2, sp = mod3(p + 1) * C::R2;
        double gl[3];
        grad19s<C::Z2>(r_lap + sm, r_lap + s0, r_lap + sp, q2, gl);
        grad19s<C::Z2>(r_pp + sm, r_pp + s0, r_pp + sp, q2, o.gp);
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
        return o.Pt - o.rho * (1. / 3.);
    };

    const int oym = (g.wy(y - 1) - y) * nz, oyp = (g.wy(y + 1) - y) * nz;
    const int ozm = g.wz(z - 1) - z, ozp = g.wz(z + 1) - z;
    const double omega = mp.omega, om1 = 1. - omega, hw = 1. - 0.5 * omega;

    SweepLocal cur, nxt;
    cur.phi = cur.rho = 1.0; cur.Fx = cur.Fy = cur.Fz = 0.0; cur.gp[0] = cur.gp[1] = cur.gp[2] = 0.0;
    cur.u0 = cur.u1 = cur.u2 = cur.Pt = 0.0;
    nxt = cur;

    // S2 cells of this thread as ring offsets (halo-3 / halo-2 coordinates); a lane past the region in a partly active
    // warp repeats its first cell, so that every round is warp-uniform
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

    // S4 of the own node: grad psi_rho, the 2 x 19 collisions, push, post-collision values back into the stage
    auto collide = [&](int x, double *sf) {
        const int q1 = (ty + 1) * C::Z1 + tz + 1;
        double ge[3];
        grad19s<C::Z1>(r_pr + ((x - 1) & 3) * C::R1, r_pr + (x & 3) * C::R1, r_pr + ((x + 1) & 3) * C::R1, q1, ge);
        const double phi = cur.phi, rho = cur.rho;
        const double u0 = cur.u0, u1 = cur.u1, u2 = cur.u2, Pt = cur.Pt;
        const double usqr = 1.5 * (u0 * u0 + u1 * u1 + u2 * u2);
        // forcing terms regrouped around Gamma_k = eqf_k / phi (hcz3d_fused.cu has the derivation)
        const double ffs = -hw * 3.0 * phi * fast_rcp(rho);
        const double D0 = hw * (cur.Fx - ge[0]), D1 = hw * (cur.Fy - ge[1]), D2 = hw * (cur.Fz - ge[2]);
        const double E0 = hw * ge[0], E1 = hw * ge[1], E2 = hw * ge[2];
        const double G0 = ffs * cur.gp[0], G1 = ffs * cur.gp[1], G2 = ffs * cur.gp[2];
        const double uD = u0 * D0 + u1 * D1 + u2 * D2;
        const double uE = u0 * E0 + u1 * E1 + u2 * E2;
        const double opg = omega * phi - (u0 * G0 + u1 * G1 + u2 * G2);
        const double rho3 = rho * (1. / 3.);
        const double Aa = (omega * (1. / 18.)) * Pt, Ba = (omega * (1. / 18.)) * rho3;
        const double Ad = (omega * (1. / 36.)) * Pt, Bd = (omega * (1. / 36.)) * rho3;
        const int xp = g.wx(x + 1), xm = g.wx(x - 1);
        const int i = (x + G) * plane + yz;
        const int oxm = (xm - x) * plane, oxp = (xp - x) * plane;
        double *const fo = P.f, *const go = P.g;
        const unsigned i0 = (unsigned)i, im = (unsigned)(i + oxm), ip = (unsigned)(i + oxp);
#pragma unroll
        for (int k = 0; k < 19; ++k) {
            const double fk = sf[k * NT];
            const double gk = sf[(19 + k) * NT];
            const double t = L19s::t(k);
            double pf, pg;
            if (k == 9) {
                const double Gam = fma(-t, usqr, t);
                pf = fma(Gam, opg, om1 * fk);
                pg = fma(om1, gk, (omega * t) * fma(-rho3, usqr, Pt)) - fma(Gam, uD, t * uE);
                if ((!TS || edge_cta) && !(ko & 16)) {
                    fo[i0 + P.kn[k]] = pf;
                    go[i0 + P.kn[k]] = pg;
                }
            } else {
                const bool axis = (L19s::cx(k) != 0) + (L19s::cy(k) != 0) + (L19s::cz(k) != 0) == 1;
                const double cu = cdot<L19s>(k, u0, u1, u2);
                const double poly = fma(cu, fma(4.5, cu, 3.0), -usqr);
                const double Gam = fma(t, poly, t);
                const double dD = cdot<L19s>(k, D0, D1, D2) - uD;
                const double dE = cdot<L19s>(k, E0, E1, E2) - uE;
                const double dG = cdot<L19s>(k, G0, G1, G2) + opg;
                pf = fma(Gam, dG, om1 * fk);
                pg = fma(t, dE, fma(Gam, dD, fma(om1, gk, fma(axis ? Ba : Bd, poly, axis ? Aa : Ad))));
                if (!(TS && L19s::cz(k) == 0) || edge_cta) {
                    unsigned idx = (L19s::cx(k) < 0 ? im : (L19s::cx(k) > 0 ? ip : i0)) + P.kn[k];
                    if (L19s::cy(k)) idx += (unsigned)(L19s::cy(k) < 0 ? oym : oyp);
                    if (L19s::cz(k)) idx += (unsigned)(L19s::cz(k) < 0 ? ozm : ozp);
                    if (!(ko & 16)) {
                        fo[idx] = pf;
                        go[idx] = pg;
                    }
                }
            }
            sf[k * NT] = pf;             // the gather of the next iteration reads these
            sf[(19 + k) * NT] = pg;
        }
    };

    // S5 for one cell: `slot` = its accumulator slot, (dy, dz) = its tile coordinates, xsrc = the plane whose pushes are in S.
    // Completes plane xsrc-1 and hands it to `store(plane, values[5])`.
    auto accumulate = [&](const PushSums &s, int slot, int xsrc, auto store) {
        double *Ts = acc + slot, *As = acc + 5 * C::NACC + slot;    // Ts[j * NACC], As[j * NACC]
        double T[5], A[4], v[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) T[j] = Ts[j * C::NACC];
#pragma unroll
        for (int j = 0; j < 4; ++j) A[j] = As[j * C::NACC];
        fold_pushes(T, A, s, xsrc == 0, v);
        if (xsrc >= 1) store(xsrc - 1, v);
        else if (wrapx) store(nx - 1, v);      // parked: completed after the march
#pragma unroll
        for (int j = 0; j < 5; ++j) Ts[j * C::NACC] = T[j];
#pragma unroll
        for (int j = 0; j < 4; ++j) As[j * C::NACC] = A[j];
    };
    auto store_own = [&](int xp, const double *v) {
        const int i = (xp + G) * plane + yz;
#pragma unroll
        for (int m = 0; m < 5; ++m) Mout.m[m][i] = v[m];
    };
    auto store_ring = [&](int xp, const double *v) {
        const size_t i = (size_t)xp * eg.eplane + ring_e;
        Mout.ephi[i] = v[0];
        double2 *q = reinterpret_cast<double2 *>(Mout.e4 + i * 4);
        q[0] = make_double2(v[1], v[2]);
        q[1] = make_double2(v[3], v[4]);
    };

    load_phi(-3);
    // x = plane being collided; the first six iterations only fill the pipeline, the last one only gathers plane nx-1
    for (int x = -6; x <= nx; ++x) {
        // ---- S1: phi of plane x+3 from the registers (edge sums added now), then prefetch the next plane's; raw moments of
        //      plane x+1 for S3 of THIS iteration, L2 prefetch of plane x+2's ----
        const bool s3_on = x + 1 >= -1 && x < nx;
        if (x < nx) {
            double *dst = r_phi + ((x + 3) & 3) * C::R3;
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (w_idx[j] >= 0) dst[w_idx[j]] = phi_n[j];
            if (w_idx[2] >= 0) dst[w_idx[2]] = ((phi_n[2] + phi_e[0]) + phi_e[1]) + phi_e[2];
            if constexpr (VAR & 2) {
                if (s3_on) load_mom(x + 1);
                if (x + 1 < nx) load_phi(x + 4);
            } else {
                if (x + 1 < nx) load_phi(x + 4);
                if (s3_on) load_mom(x + 1);
            }
        }
        __syncthreads();
        if constexpr (TS) {
            // push of plane x-1: its post-collision tile of every direction, shifted by c_k, straight from the stage
            if (tid == TMA_TID && x >= 1 && !(ko & 1024) && !edge_cta) {
                const int xq = x - 1;
                const uint32_t src = stage_a + (xq & 1) * C::STAGE_BYTES;
                const int xs0 = xq + G, xsm = g.wx(xq - 1) + G, xsp = g.wx(xq + 1) + G;
#pragma unroll
                for (int k = 0; k < 19; ++k) {
                    if (L19s::cz(k) != 0) continue;      // push_by_box: the rest of the rule is edge_cta
                    const int xs = L19s::cx(k) < 0 ? xsm : (L19s::cx(k) > 0 ? xsp : xs0);
                    const PushBox b = push_box_start<TY, TZ>(y0, z0, L19s::cy(k));
                    tma_store_4d(&tmap_fo, src + k * NT * 8, b.z, b.y, xs, k);
                    tma_store_4d(&tmap_go, src + C::SET_BYTES + k * NT * 8, b.z, b.y, xs, k);
                }
                bulk_commit();
            }
        }

        // ---- S5: what plane x-1 pushed (its post-collision values are in its stage) ----
        if (x >= 1 && !(ko & 2)) {
            const double *S = reinterpret_cast<const double *>(smem_raw + ((x - 1) & 1) * C::STAGE_BYTES);
            PushSums ps;
            gather_pushes<TY, TZ>(S, ty, tz, ps);
            accumulate(ps, tid, x - 1, store_own);
            if (h_act && !(ko & 1)) {
                // ring cell: only the directions leaving the tile through that side can contribute
                const int dy = h1y - 1, dz = h1z - 1;
                if (tid < C::Z1) gather_pushes<TY, TZ, -1, 2>(S, dy, dz, ps);              // row below the tile
                else if (tid < 2 * C::Z1) gather_pushes<TY, TZ, 1, 2>(S, dy, dz, ps);      // row above
                else if ((tid - 2 * C::Z1) & 1) gather_pushes<TY, TZ, 2, 1>(S, dy, dz, ps);   // column behind the last one
                else gather_pushes<TY, TZ, 2, -1>(S, dy, dz, ps);                          // column before the first one
                accumulate(ps, NT + tid, x - 1, store_ring);
            }
        }
        if (s3_on) merge_mom();   // the loads of S1 have had the whole gather phase to land (merging after S2 instead: no change, 21.75 vs 21.72 ms)
        // ---- S2: lap(phi), psi(phi) of plane x+2 on tile + halo 2 ----
        if (x + 2 >= -2 && x < nx && !(ko & 32)) {
            const int p = x + 2;
            const double *Pm = r_phi + ((p - 1) & 3) * C::R3, *P0 = r_phi + (p & 3) * C::R3, *Pp = r_phi + ((p + 1) & 3) * C::R3;
            double lap_v[C::N2], pp_v[C::N2];
#pragma unroll
            for (int j = 0; j < C::N2; ++j) {
                if (!c2_on[j]) continue;
                const int q3 = c2_q3[j];
                const double phi_c = P0[q3];
                double sa[2] = {0.0, 0.0}, sd[2] = {0.0, 0.0};
                int na = 0, nd = 0;
#pragma unroll
                for (int k = 0; k < 19; ++k) {
                    if (k == L19s::REST) continue;
                    const double *R = L19s::cx(k) < 0 ? Pm : (L19s::cx(k) > 0 ? Pp : P0);
                    const double v = R[q3 + L19s::cy(k) * C::Z3 + L19s::cz(k)];
                    const bool axis = (L19s::cx(k) != 0) + (L19s::cy(k) != 0) + (L19s::cz(k) != 0) == 1;
                    if (axis) { sa[na & 1] += v; ++na; }
                    else { sd[nd & 1] += v; ++nd; }
                }
                lap_v[j] = 6.0 * ((1. / 18.) * (sa[0] + sa[1]) + (1. / 36.) * (sd[0] + sd[1]) - (2. / 3.) * phi_c);
                pp_v[j] = hcz_psi1(phi_c, mp.a, mp.b);
            }
            const int sl = mod3(p) * C::R2;
#pragma unroll
            for (int j = 0; j < C::N2; ++j) {
                if (!c2_on[j]) continue;
                r_lap[sl + c2_q2[j]] = lap_v[j];
                r_pp[sl + c2_q2[j]] = pp_v[j];
            }
        }
        __syncthreads();
        // the stage of plane x-1 has been gathered: refill it with plane x+1 (generic-proxy accesses before the async write)
        if (tid == TMA_TID && x >= 1 && x + 1 < nx) {
            if constexpr (TS) bulk_wait_read_all();    // the box stores of plane x-1 have read the stage
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(x + 1);
        }

        // ---- S3: level 2 of plane x+1 on tile + halo 1 ----
        if (s3_on && !(ko & 256)) {
            double *pr = r_pr + ((x + 1) & 3) * C::R1;
            if (h_warp && !(ko & 4)) {
                SweepLocal tmp;
                const double a = level2(x + 1, ty + 1, tz + 1, mo_c, nxt);
                const double b = level2(x + 1, h1y, h1z, mh_c, tmp);
                pr[(ty + 1) * C::Z1 + tz + 1] = a;
                pr[h1y * C::Z1 + h1z] = b;
            } else {
                pr[(ty + 1) * C::Z1 + tz + 1] = level2(x + 1, ty + 1, tz + 1, mo_c, nxt);
            }
        }
        __syncthreads();

        // ---- S4: collide + push plane x ----
        if (x >= 0 && x < nx) {
            mbar_wait(&mbar[x & 1], (x >> 1) & 1);
            if (!(ko & 128)) collide(x, reinterpret_cast<double *>(smem_raw + (x & 1) * C::STAGE_BYTES) + tid);
            if constexpr (TS) fence_proxy_async_smem();   // the stage as the box stores (async proxy) must see it, before the next barrier
        }
        cur = nxt;
    }

    if constexpr (TS) {
        if (tid == TMA_TID) bulk_wait_all();
    }
    // ---- x wraps inside the CTA: plane nx-1 still lacks the C group of plane 0 (parked in its slot at the start), plane 0
    //      the A group of plane nx-1 (in the accumulators now).  Same thread wrote those slots: a plain read-modify-write. ----
    if (wrapx) {
        auto finish = [&](int slot, auto ref) {   // ref(m, last) -> reference to moment m of the cell in plane nx-1 (last) / plane 0
            const double *Ts = acc + slot, *As = acc + 5 * C::NACC + slot;
            double T[5], A[4];
#pragma unroll
            for (int j = 0; j < 5; ++j) T[j] = Ts[j * C::NACC];
#pragma unroll
            for (int j = 0; j < 4; ++j) A[j] = As[j * C::NACC];
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                double &l = ref(m, true), &f = ref(m, false);
                l = finish_last(T, m, l);
                f = finish_first(A, m, f);
            }
        };
        finish(tid, [&](int m, bool last) -> double & { return Mout.m[m][(size_t)((last ? nx - 1 : 0) + G) * plane + yz]; });
        if (h_act)
            finish(NT + tid, [&](int m, bool last) -> double & {
                const size_t i = (size_t)(last ? nx - 1 : 0) * eg.eplane + ring_e;
                return m == 0 ? Mout.ephi[i] : Mout.e4[i * 4 + (m - 1)];
            });
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------
static constexpr int SW_TY = 8, SW_TZ = 32;
static constexpr int SW_VAR_DEFAULT = 5;   // CLBM_HCZ3D_SWEEP_VAR overrides (kernel header)

EdgeGeom hcz3d_sweep_edge_geom(const clbm_ctx *c) { return make_edge_geom<SW_TY, SW_TZ>(c->geo.ny, c->geo.nz); }
long long hcz3d_sweep_edge_doubles(const clbm_ctx *c) { return (long long)c->geo.nx * hcz3d_sweep_edge_geom(c).eplane; }

// geometry the sweep kernel can run on (the caller also needs a wall-free mask and the default fused path)
bool hcz3d_sweep_shape_ok(const clbm_ctx *c)
{
    const Geom &g = c->geo;
    return c->prm.model == CLBM_MODEL_HCZ_D3Q19 && g.ny % SW_TY == 0 && g.nz % SW_TZ == 0 && g.nx >= 4 && g.ny >= SW_TY && g.nz >= SW_TZ &&
           g.ncs < (1LL << 31) && 19ull * (unsigned long long)g.ncs < (1ull << 32) && (long long)g.nx * make_edge_geom<SW_TY, SW_TZ>(g.ny, g.nz).eplane < (1LL << 31) && get_encode() != nullptr;
}

// one sweep: populations pop[*][parity] -> pop[*][1 - parity], moments mom[src] (+ edges) -> mom[1 - src] (+ edges)
int hcz3d_sweep_launch(clbm_ctx *c, int src)
{
    using C = SweepCfg<SW_TY, SW_TZ>;
    const Geom &g = c->geo;
    CUtensorMap tm[2];
    const cuuint32_t box[4] = {(cuuint32_t)SW_TZ, (cuuint32_t)SW_TY, 1, 19};
    for (int s = 0; s < 2; ++s)
        if (int rc = cached_tmap(c, c->pop[s][c->parity], box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, &tm[s])) return rc;
    SweepOut P = {c->pop[0][1 - c->parity], c->pop[1][1 - c->parity], {0}};
    for (int k = 0; k < 19; ++k) P.kn[k] = (unsigned)((unsigned long long)k * (unsigned long long)g.ncs);
    SweepMom Min, Mout;
    for (int m = 0; m < 5; ++m) {
        Min.m[m] = c->mom[src][m];
        Mout.m[m] = c->mom[1 - src][m];
    }
    Min.ephi = c->mome[src][0];
    Min.e4 = c->mome[src][1];
    Mout.ephi = c->mome[1 - src][0];
    Mout.e4 = c->mome[1 - src][1];
    const int var_env = c->env.hcz3d_sweep_var >= 0 ? c->env.hcz3d_sweep_var : SW_VAR_DEFAULT;
    const int ko = c->env.hcz3d_sweep_ko > 0 ? c->env.hcz3d_sweep_ko : 0;
    const int var = ko ? 3 : ((var_env & 4) ? 2 : (var_env & 1));       // kernel table slot: VAR 0, 1, 5, and 5 with the knock-out tests compiled in
    using Kern = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const SweepOut, const SweepMom, const SweepMom,
                          Geom, ModelParams, EdgeGeom, int);
    static const Kern kerns[4] = {hcz3d_sweep_kernel<SW_TY, SW_TZ, 0>, hcz3d_sweep_kernel<SW_TY, SW_TZ, 1>, hcz3d_sweep_kernel<SW_TY, SW_TZ, 5>,
                                  hcz3d_sweep_kernel<SW_TY, SW_TZ, 5, true>};
    const Kern kern = kerns[var];
    static PerDeviceOnce attr[4];
    if (attr[var].need(c->device)) {
        CLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr[var].mark(c->device);
    }
    // output boxes of the TMA-store form: one direction's tile
    CUtensorMap tmo[2];
    const cuuint32_t obox[4] = {(cuuint32_t)SW_TZ, (cuuint32_t)SW_TY, 1, 1};
    for (int s = 0; s < 2; ++s)
        if (int rc = cached_tmap(c, c->pop[s][1 - c->parity], obox, CU_TENSOR_MAP_L2_PROMOTION_NONE, &tmo[s])) return rc;
    dim3 grid(g.nz / SW_TZ, g.ny / SW_TY, 1);
    LaunchScope ls(c, "hcz3d_sweep_collide_stream_moments", true);
    kern<<<grid, C::NT, C::SMEM, c->stream>>>(tm[0], tm[1], tmo[0], tmo[1], P, Min, Mout, g, c->mp, hcz3d_sweep_edge_geom(c), ko);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}


// edge sums of the planes [x0, x0 + np) of set `set` <- 0 (those planes' node arrays hold complete sums)
__global__ void __launch_bounds__(256) zero_edges_kernel(double *ephi, double *e4, long long ep, int xa, int xb, int np)
{
    // blockIdx.y: 0 = planes [xa, xa + np), 1 = planes [xb, xb + np)
    const long long x0 = blockIdx.y ? xb : xa;
    const long long n1 = ep * np, n4 = 4 * ep * np;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        if (i < n1) ephi[ep * x0 + i] = 0.0;
        e4[4 * ep * x0 + i] = 0.0;
    }
}

int hcz3d_sweep_zero_edges(clbm_ctx *c, int set, int x0, int np)
{
    const size_t ep = (size_t)hcz3d_sweep_edge_geom(c).eplane;
    CLBM_CUDA(cudaMemsetAsync(c->mome[set][0] + ep * x0, 0, ep * np * sizeof(double), c->stream));
    CLBM_CUDA(cudaMemsetAsync(c->mome[set][1] + 4 * ep * x0, 0, 4 * ep * np * sizeof(double), c->stream));
    return 0;
}

// the same for the two boundary planes of a slab in one launch (stage 0 of every slab step)
int hcz3d_sweep_zero_edge_planes(clbm_ctx *c, int set, int xa, int xb)
{
    const long long ep = hcz3d_sweep_edge_geom(c).eplane;
    LaunchScope ls(c, "zero_edge_planes");
    zero_edges_kernel<<<dim3(64, 2), 256, 0, c->stream>>>(c->mome[set][0], c->mome[set][1], ep, xa, xb, 1);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

// phi of the planes [x0, x0 + np) with the edge sums folded in, densely into dst (the moment-halo pack of an x-slab)
__global__ void __launch_bounds__(256)
pack_phi_merged_kernel(const double *__restrict__ M, const double *__restrict__ E, double *__restrict__ dst, Geom g, EdgeGeom eg, int x0, long long n)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int xp = x0 + (int)(t / g.plane), r = (int)(t % g.plane);
    const int yy = r / g.nz, zz = r % g.nz;
    int e[3];
    edge_offsets<SW_TY, SW_TZ>(eg, yy, zz, e);
    const double *Ep = E + (size_t)xp * eg.eplane;
    const double e0 = e[0] >= 0 ? Ep[e[0]] : 0.0, e1 = e[1] >= 0 ? Ep[e[1]] : 0.0, e2 = e[2] >= 0 ? Ep[e[2]] : 0.0;
    dst[t] = ((M[(size_t)(xp + g.G) * g.plane + r] + e0) + e1) + e2;
}

int hcz3d_pack_phi_merged(clbm_ctx *c, double *dst, int x0, int nplanes)
{
    const long long n = (long long)nplanes * c->geo.plane;
    LaunchScope ls(c, "pack_phi_merged");
    pack_phi_merged_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(c->mom[c->mom_src][0], c->mome[c->mom_src][0], dst, c->geo, hcz3d_sweep_edge_geom(c), x0, n);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace clbm
