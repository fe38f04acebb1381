// hcz2d_fused.cu -- He-Chen-Zhang D2Q9 time step as ONE column-marching kernel (PF/apps/rayleighTaylor2D.h).
//
// The staged path (hcz2d_kernels.cu) reads f twice and round-trips five scalar fields through HBM
// (~410 B per lattice update against 289 algorithmic).  Here a CTA owns a segment of the y axis (y is the
// fastest index, so a warp touches 256 contiguous bytes of every population array) and marches along x:
//
//   column x+2 : phi = sum_k f_k, then rho(phi), psi(phi), psi(rho)            -> 5-slot shared-memory rings
//   column x+1 : lap(phi) with the wall mirror rule (:467-495)                 -> ring
//   column x   : grad lap phi, grad psi(phi), grad psi(rho), grad rho from the rings (:341-446, :501-529),
//                velocity (:316-337), total_P (:452-460), collideBgk of f and g (:552-606, rest :642-663),
//                push stream with half-way bounce-back (:533-549)
//
// Every population is fetched from HBM once (the second read of f at collide time is two columns behind its
// first read and hits L1/L2) and written once; no scalar field leaves the SM.
//
// Segments overlap by the stencil reach instead of having dedicated halo threads: a CTA of NT threads covers
// NT consecutive rows, computes phi on all of them, lap(phi) on rows 1..NT-2 and collides rows 2..NT-3.
// x-slab mode: phi of the ghost columns (-2,-1,nx,nx+1) comes from the exchanged moment halo (fld[0]).
//
// The per-node arithmetic is the staged kernel's with the divisions by constants turned into multiplications and the
// quotients shared (4 FP64 divisions per node instead of 13; they were a third of the instruction stream): fused and
// staged agree to O(1 ulp) per step, both within 1e-10 of the oracle.
#include <cooperative_groups.h>

#include <cstdlib>
#include <type_traits>

#include "mrt.cuh"
#include "sc_cell.cuh"

namespace clbm {

using L9f = D2Q9;

struct Hcz2dTables {
    const double *fin[9];
    const double *gin[9];
    double *fout[9];
    double *gout[9];
};

constexpr int HCZ2D_NS = 5;   // ring slots (columns x-2 .. x+2 are live at once)

// k-th neighbour of ring position p with the mirror rule (a bounce_back neighbour is replaced by the opposite one)
template <int NT>
CLBM_D double ring_mirror(const double (*R)[NT], const uint8_t (*FL)[NT], int k, int sm, int s0, int sp, int p)
{
    const int slot = L9f::cx(k) < 0 ? sm : (L9f::cx(k) > 0 ? sp : s0);
    if (FL[slot][p + L9f::cy(k)] == CELL_BB) {
        const int oslot = L9f::cx(k) < 0 ? sp : (L9f::cx(k) > 0 ? sm : s0);
        return R[oslot][p - L9f::cy(k)];
    }
    return R[slot][p + L9f::cy(k)];
}

template <int NT>
CLBM_D void ring_grad(const double (*R)[NT], const uint8_t (*FL)[NT], unsigned wall, int sm, int s0, int sp, int p,
                      double &gx, double &gy)
{
    double ax = 0.0, ay = 0.0;
    if (wall == 0u) {
        // opposite directions first (equal weights): half the FMAs and chains of three instead of six
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ko = L9f::opp(k);
            const int slot = L9f::cx(k) < 0 ? sm : (L9f::cx(k) > 0 ? sp : s0);
            const int oslot = L9f::cx(ko) < 0 ? sm : (L9f::cx(ko) > 0 ? sp : s0);
            const double d = R[slot][p + L9f::cy(k)] - R[oslot][p + L9f::cy(ko)];
            if (L9f::cx(k)) ax += (L9f::t(k) * L9f::cx(k)) * d;
            if (L9f::cy(k)) ay += (L9f::t(k) * L9f::cy(k)) * d;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            if (k == L9f::REST) continue;
            const double v = ring_mirror<NT>(R, FL, k, sm, s0, sp, p);
            if (L9f::cx(k)) ax += L9f::t(k) * L9f::cx(k) * v;
            if (L9f::cy(k)) ay += L9f::t(k) * L9f::cy(k) * v;
        }
    }
    gx = 3.0 * ax;
    gy = 3.0 * ay;
}

// MRT = true: CLBM_COLLISION_MRT (include/clbm.h, mrt.cuh) -- a compile-time variant of the collide phase only; the BGK
// instantiations are unchanged.
//
// MULTI = true: L2-resident lattices (BASELINE configs[1], 256 x 1026: 75.6 MB for both buffers), `nsteps` time steps in ONE
// cooperative launch.  Launch by launch such a lattice is latency bound and half of a CTA's work is its prologue: a 4-column chunk
// reads the populations of 8 columns for phi.  Here a step is two phases separated by grid barriers:
//   A  every CTA sums phi of ITS OWN nodes into a scalar field (9 loads + 1 store per node, L2 to L2);
//   B  the march above with phi of EVERY column taken from that field (the path the ghost columns of an x-slab take): the
//      prologue is 4 doubles per thread instead of 36, the populations are read once, all of them issued at the top of the column
//      iteration (registers) so that the ring phases and barriers of the iteration cover their L2 latency.
// The buffers swap roles every step.  Same cell functions, same summation order; the compiler contracts a few products of the BGK
// collide differently in this instantiation, so the populations agree with the launch-per-step form to O(1 ulp) per step (MRT:
// bit for bit) -- tests/test_gpu_zq_hcz2d_multistep.py.
// MEASURED (tools/hcz2d_multi.py, B200): 256 x 1026 31.3 us per step against 24.6 launch by launch, MRT 36.7 against 30.8,
// 128 x 514 15.8 against 18.4.  A column iteration costs ~3.4 us whoever issues it (16 warps per SM, FP64 dependency chains: the
// rate the kernel also runs at from HBM, 17.9 cycles per node and SM, puts configs[1] at 16 us per step at best), so the phi pass and
// the two grid barriers cost more than the prologue they save.  The form is therefore OPT-IN (CLBM_HCZ2D_MULTI=1), not a default.
template <int NT, int MINB, bool MRT = false, bool MULTI = false>
__global__ void __launch_bounds__(NT, MINB)
hcz2d_fused_kernel(const Hcz2dTables P, const uint8_t *__restrict__ flag, const double *__restrict__ phi_g, Geom g,
                   ModelParams mp, int xchunk, int x_begin, int x_end, int nsteps, double *phi_w)
{
    constexpr int NS = HCZ2D_NS;
    __shared__ double r_phi[NS][NT], r_rho[NS][NT], r_pp[NS][NT], r_pr[NS][NT], r_lap[NS][NT];
    __shared__ uint8_t r_fl[NS][NT];
    // the 9 g populations the collision of column x consumes are copied global -> shared one column ahead with cp.async
    // (no registers, no stall at the point of use: long_scoreboard was the second stall reason); a thread only ever reads
    // the slots it filled itself, so the copies need no barrier, only cp.async.wait_group
    __shared__ double st_g[2][9][NT];     // g only: the second read of f hits L1/L2 (it was read two columns earlier for phi)
    extern __shared__ double st_f_raw[];          // MULTI reads f once: [9][NT] doubles of dynamic shared memory, staged at the top of
    double (*st_f)[NT] = reinterpret_cast<double (*)[NT]>(st_f_raw);   // its own column's iteration (static + dynamic > 48 KB)

    const int tid = threadIdx.x;
    const int ny = g.ny, G = g.G;
    const int yy = (int)blockIdx.x * (NT - 4) - 2 + tid;   // unwrapped row of this thread
    const int yw = g.wy(yy);
    const bool has_phi = yy <= ny + 1;
    const bool has_lap = tid >= 1 && tid < NT - 1 && yy <= ny;
    const bool own = tid >= 2 && tid < NT - 2 && yy < ny;
    // columns [x_begin, x_end) of the slab: the whole slab, or one of the ranges of the overlap protocol (boundary columns /
    // interior) -- a column is collided by exactly one launch, and a (node, direction) slot is written by exactly one column
    const int xa = x_begin + blockIdx.y * xchunk;
    const int xb = min(x_end, xa + xchunk);

    // MULTI: the buffers swap roles every step, so the four population sets are addressed as base + k * ncs (four pointers in
    // registers) instead of through the 36 pointers of the parameter block
    const double *fin_b = P.fin[0], *gin_b = P.gin[0];
    double *fout_b = P.fout[0], *gout_b = P.gout[0];
    const size_t ncs_ = (size_t)g.ncs;
    auto FIN = [&](int k) -> const double * { if constexpr (MULTI) return fin_b + k * ncs_; else return P.fin[k]; };
    auto GIN = [&](int k) -> const double * { if constexpr (MULTI) return gin_b + k * ncs_; else return P.gin[k]; };
    auto FOUT = [&](int k) -> double * { if constexpr (MULTI) return fout_b + k * ncs_; else return P.fout[k]; };
    auto GOUT = [&](int k) -> double * { if constexpr (MULTI) return gout_b + k * ncs_; else return P.gout[k]; };

    auto slot_of = [](int xg) { return (xg + 2 * NS) % NS; };
    auto col_of = [&](int xg) { return (g.wx(xg) + G) * ny + yw; };   // storage index of (xg, this row)
    auto is_ghost = [&](int xg) { return !g.wrapx && (xg < 0 || xg >= g.nx); };

    // phi (and the scalars that depend on phi alone) of column xg into the rings
    auto put_phi = [&](int xg, double phi, uint8_t fl) {
        const int s = slot_of(xg);
        const double rho = mp.rho_g + ((phi - mp.phi_g) * mp.inv_dphi) * mp.drho;
        r_phi[s][tid] = phi;
        r_rho[s][tid] = rho;
        r_pp[s][tid] = hcz_psi1(phi, mp.a, mp.b);
        r_pr[s][tid] = hcz_psi1(rho, mp.a, mp.b);
        r_fl[s][tid] = fl;
    };
    auto load_f = [&](int xg, double *f) {
        const int i = col_of(xg);
#pragma unroll
        for (int k = 0; k < 9; ++k) f[k] = MULTI ? __ldcg(FIN(k) + i) : FIN(k)[i];
    };
    auto fill_direct = [&](int xg) -> bool {   // true: the cell is a bounce_back node
        if (!has_phi) return false;
        const int i = col_of(xg);
        const uint8_t fl = flag[i];
        if constexpr (MULTI) { put_phi(xg, __ldcg(phi_w + i), fl); return fl == CELL_BB; }
        if (is_ghost(xg)) { put_phi(xg, phi_g[i], fl); return fl == CELL_BB; }
        double f[9];
        load_f(xg, f);
        put_phi(xg, Mom<L9f>::sum(f), fl);
        return fl == CELL_BB;
    };
    // Walls are rare (two rows of the whole lattice in the shipped cases), so every stencil phase exists twice: the general
    // one with the mirror rule per neighbour, and a straight-line one for windows without a bounce_back node.  The choice
    // is CTA-uniform: bit (xg & 7) of wmask says "column xg has a bounce_back node on this CTA's rows".
    unsigned wmask = 0;
    auto walls_near = [&](int xg) { return (wmask & ((1u << ((xg - 1) & 7)) | (1u << (xg & 7)) | (1u << ((xg + 1) & 7)))) != 0u; };
    auto make_lap = [&](int xg) {
        if (!has_lap) return;
        const int s0 = slot_of(xg), sm = slot_of(xg - 1), sp = slot_of(xg + 1);
        double sum = 0.0;
        if (walls_near(xg)) {
            if (r_fl[s0][tid] == CELL_BULK) {
                const double phi_c = r_phi[s0][tid];
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    if (k != L9f::REST) sum += L9f::t(k) * (ring_mirror<NT>(r_phi, r_fl, k, sm, s0, sp, tid) - phi_c);
            }
        } else {
            // sum_k t_k (phi_nb - phi_c) = (1/9) S_axis + (1/36) S_diag - (5/9) phi_c
            const double sa = (r_phi[sm][tid] + r_phi[sp][tid]) + (r_phi[s0][tid - 1] + r_phi[s0][tid + 1]);
            const double sd = (r_phi[sm][tid - 1] + r_phi[sp][tid + 1]) + (r_phi[sm][tid + 1] + r_phi[sp][tid - 1]);
            sum = fma(1. / 9., sa, fma(1. / 36., sd, (-5. / 9.) * r_phi[s0][tid]));
        }
        r_lap[s0][tid] = 6.0 * sum;
    };

  for (int step = 0; step < (MULTI ? nsteps : 1); ++step) {
    if constexpr (MULTI) {
        if (step & 1) {
            fin_b = P.fout[0]; gin_b = P.gout[0];
            fout_b = const_cast<double *>(P.fin[0]); gout_b = const_cast<double *>(P.gin[0]);
        } else {
            fin_b = P.fin[0]; gin_b = P.gin[0];
            fout_b = P.fout[0]; gout_b = P.gout[0];
        }
        // ---- phase A: phi of the own nodes of this CTA's columns into the scalar field ----
        if (own) {
            for (int x0 = xa; x0 < xb; x0 += 2) {      // two columns at a time: 18 loads in flight (four spilled)
                double fa[2][9];
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    if (x0 + j < xb) load_f(x0 + j, fa[j]);
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    if (x0 + j < xb) __stcg(phi_w + col_of(x0 + j), Mom<L9f>::sum(fa[j]));
            }
        }
        cooperative_groups::this_grid().sync();
    }
    // ---- prologue: phi of columns xa-2 .. xa+1, lap of columns xa-1, xa ----
    int pw = fill_direct(xa - 2);
    pw |= (int)fill_direct(xa - 1);
    pw |= (int)fill_direct(xa);
    pw |= (int)fill_direct(xa + 1);
    double fn[9];
    double phin = 0.0;    // MULTI: phi of the column after next, fetched one iteration ahead like fn
    uint8_t fln = CELL_BULK;
    bool ghost_n = is_ghost(xa + 2);
    if constexpr (MULTI) {
        if (has_phi) { const int i = col_of(xa + 2); phin = __ldcg(phi_w + i); fln = flag[i]; }
    } else {
        if (has_phi && !ghost_n) { load_f(xa + 2, fn); fln = flag[col_of(xa + 2)]; }
    }
    wmask = __syncthreads_or(pw) ? 0xffu : 0u;   // conservative for the four prologue columns; the ring corrects itself as it advances
    make_lap(xa - 1);
    make_lap(xa);

    const double omega = mp.omega, hw = 1. - 0.5 * omega;
    const int oym = (g.wy(yy - 1) - yy), oyp = (g.wy(yy + 1) - yy);
    auto stage_pops = [&](int xg) {     // own row of column xg (always a real column of this slab) into slot xg & 1
        if (own) {
            const int i = (xg + G) * ny + yy;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(&st_g[xg & 1][k][tid])), "l"(GIN(k) + i) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage_pops(xa);

    for (int x = xa; x < xb; ++x) {
        if constexpr (MULTI) {      // f of this column rides in the group of the next column's g; the collide waits for both
            if (own) {
                const int i = (x + G) * ny + yy;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(&st_f[k][tid])), "l"(FIN(k) + i) : "memory");
                }
            }
        }
        if (x + 1 < xb) stage_pops(x + 1);
        else asm volatile("cp.async.commit_group;" ::: "memory");   // keep one group per iteration so wait_group 1 means "column x is here"
        // 1. column x+2: phi from the prefetched populations (or the exchanged ghost field)
        int anyw = 0;
        if (has_phi) {
            if constexpr (MULTI) put_phi(x + 2, phin, fln);
            else if (ghost_n) { const int i = col_of(x + 2); fln = flag[i]; put_phi(x + 2, phi_g[i], fln); }
            else put_phi(x + 2, Mom<L9f>::sum(fn), fln);
            anyw = fln == CELL_BB;
        }
        // prefetch column x+3 for the next iteration
        ghost_n = is_ghost(x + 3);
        if constexpr (MULTI) {
            if (x + 1 < xb && has_phi) { const int i = col_of(x + 3); phin = __ldcg(phi_w + i); fln = flag[i]; }
        } else {
            if (x + 1 < xb && has_phi && !ghost_n) { load_f(x + 3, fn); fln = flag[col_of(x + 3)]; }
        }
        {
            const unsigned bit = 1u << ((x + 2) & 7);
            wmask = __syncthreads_or(anyw) ? (wmask | bit) : (wmask & ~bit);
        }
        // 2. column x+1: lap(phi)
        make_lap(x + 1);
        __syncthreads();
        // 3. column x: collide + push
        const int s0 = slot_of(x), sm = slot_of(x - 1), sp = slot_of(x + 1);
        if (!own || r_fl[s0][tid] != CELL_BULK) continue;

        auto collide = [&](auto wtag) {
        constexpr bool W = decltype(wtag)::value;
        const int i = (x + G) * ny + yy;
        double f[9], gg[9];
        if constexpr (MULTI) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 9; ++k) { gg[k] = st_g[x & 1][k][tid]; f[k] = st_f[k][tid]; }
        } else {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 9; ++k) { gg[k] = st_g[x & 1][k][tid]; f[k] = FIN(k)[i]; }
        }

        unsigned wall = 0;
        if constexpr (W) {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                if (k == 4) continue;
                const int slot = L9f::cx(k) < 0 ? sm : (L9f::cx(k) > 0 ? sp : s0);
                if (r_fl[slot][tid + L9f::cy(k)] == CELL_BB) wall |= 1u << k;
            }
        }
        double glx, gly, grx, gry, Ex, Ey, gpx, gpy;
        ring_grad<NT>(r_lap, r_fl, wall, sm, s0, sp, tid, glx, gly);
        ring_grad<NT>(r_rho, r_fl, wall, sm, s0, sp, tid, grx, gry);
        ring_grad<NT>(r_pr, r_fl, wall, sm, s0, sp, tid, Ex, Ey);
        ring_grad<NT>(r_pp, r_fl, wall, sm, s0, sp, tid, gpx, gpy);

        const double phi = r_phi[s0][tid], rho = r_rho[s0][tid];
        double jx, jy, jz;
        Mom<L9f>::first(gg, jx, jy, jz);
        const double Pt = Mom<L9f>::sum(gg);
        double forcex, forcey;
        hcz2d_force(mp, rho, glx, gly, forcex, forcey);
        // one division per node here (3/rho); every division by a constant is a multiplication by its reciprocal
        const double inv_r3 = 3.0 * fast_rcp(rho), rho3 = rho * (1.0 / 3.0);
        const double u0 = (jx + forcex * (1.0 / 6.0)) * inv_r3;
        const double u1 = (jy + forcey * (1.0 / 6.0)) * inv_r3;
        const double Pp = Pt - 0.5 * ((u0 * -grx + u1 * -gry) * (1.0 / 3.0));
        const double usqr = 1.5 * (u0 * u0 + u1 * u1);

        const int xp = g.wx(x + 1), xm = g.wx(x - 1);
        const int oxm = (xm - x) * ny, oxp = (xp - x) * ny;
        // eqf / phi = t_k (1 + poly) =: Gamma, so neither eqf nor 1/phi is formed; opposite directions share c.u, c.F, c.E,
        // c.G up to the sign and are collided in pairs (even part once, odd part added / subtracted).
        // Forcing vectors as the reference writes them: F (forcex, forcey), -E (grad psi(rho)), -grad psi(phi).
        // The two forcing terms regrouped around Gamma_k (reference form: fg = hw [ (e-u).F Gam + (e-u).(-E) (Gam - t) ],
        // ff = hw (e-u).(-grad psi(phi)) 3 Gam):
        //   fg_k = Gamma_k (c_k - u).D + t_k (c_k - u).Eh,   ff_k = Gamma_k (c_k - u).Gv
        // with the per-node vectors D = hw (F - E), Eh = hw E, Gv = -3 hw grad psi(phi); everything is written as explicit
        // FMAs (the kernel is FP64-issue bound and the compiler does not re-associate)
        const double uF = u0 * forcex + u1 * forcey, uE = u0 * Ex + u1 * Ey, uG = u0 * gpx + u1 * gpy;
        const double om1 = 1. - omega, op = omega * phi, hw3 = 3.0 * hw;
        const double D0 = hw * (forcex - Ex), D1 = hw * (forcey - Ey);
        const double E0 = hw * Ex, E1 = hw * Ey;
        const double G0 = -hw3 * gpx, G1 = -hw3 * gpy;
        const double uD = u0 * D0 + u1 * D1, uEh = u0 * E0 + u1 * E1;
        const double opg = op - (u0 * G0 + u1 * G1);                   // omega phi - u.Gv
        const double orho3 = omega * rho3;
        // omega t_k (Pp + rho/3 ev_k) = A + B ev_k with one (A, B) pair per weight class
        const double Aa = (omega * (1. / 9.)) * Pp, Ba = (omega * (1. / 9.)) * rho3;
        const double Ad = (omega * (1. / 36.)) * Pp, Bd = (omega * (1. / 36.)) * rho3;
        auto push = [&](int k, double pf, double pg) {
            if (W && (wall & (1u << k))) {
                FOUT(L9f::opp(k))[i] = pf;
                GOUT(L9f::opp(k))[i] = pg;
            } else {
                const int off = (L9f::cx(k) < 0 ? oxm : (L9f::cx(k) > 0 ? oxp : 0)) + (L9f::cy(k) < 0 ? oym : (L9f::cy(k) > 0 ? oyp : 0));
                FOUT(k)[i + off] = pf;
                GOUT(k)[i + off] = pg;
            }
        };
        if constexpr (MRT) {
            // out = in + F - M^-1 S M (in - eq + F/2), F = the forcing terms below without their (1 - omega/2) factor; with
            // Gamma_k = eqf_k / phi = t_k (1 + poly_k):  Ff_k = 3 Gamma_k (c_k - u).(-grad psi(phi)),
            // Fg_k = Gamma_k (c_k - u).F + (Gamma_k - t_k) (c_k - u).(-E); rest population as the reference writes it
            // (u.(-E), layered variant's own force: SURVEY.md B.8, B.9).  One population set at a time (registers).
            const MrtRates S = {omega, mp.s_e, mp.s_eps, mp.s_q, omega};
            double uF0 = uF;
            if (mp.sc_force == CLBM_HCZ_FORCE_LAYERED) {
                const double slope = mp.drho * mp.inv_dphi;
                double fx0, fy0;
                hcz2d_force(mp, rho, slope * glx, slope * gly, fx0, fy0);
                uF0 = u0 * fx0 + u1 * fy0;
            }
            double sv[9], vv[9], wv[9], of[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const double cu = cdot<L9f>(k, u0, u1, 0.0);
                const double Gam = L9f::t(k) * (1. + (3.0 * cu + 4.5 * cu * cu - usqr));
                const double Ff = (k == 4) ? 3.0 * uG * Gam : -3.0 * (cdot<L9f>(k, gpx, gpy, 0.0) - uG) * Gam;
                sv[k] = f[k] + Ff;
                vv[k] = f[k] - phi * Gam + 0.5 * Ff;
            }
            mrt9_relax(vv, S, wv);
#pragma unroll
            for (int k = 0; k < 9; ++k) of[k] = sv[k] - wv[k];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const double t = L9f::t(k);
                const double cu = cdot<L9f>(k, u0, u1, 0.0);
                const double poly = 3.0 * cu + 4.5 * cu * cu - usqr;
                const double Gam = t * (1. + poly);
                const double eqg = t * (Pp + rho3 * poly);
                const double Fg = (k == 4) ? (-uF0 * Gam - uE * (Gam - t))
                                           : (Gam * (cdot<L9f>(k, forcex, forcey, 0.0) - uF) - (Gam - t) * (cdot<L9f>(k, Ex, Ey, 0.0) - uE));
                sv[k] = gg[k] + Fg;
                vv[k] = gg[k] - eqg + 0.5 * Fg;
            }
            mrt9_relax(vv, S, wv);
            FOUT(4)[i] = of[4];
            GOUT(4)[i] = sv[4] - wv[4];
#pragma unroll
            for (int k = 0; k < 9; ++k)
                if (k != 4) push(k, of[k], sv[k] - wv[k]);
            return;
        }
        {   // rest population (:642-663)
            const double t = L9f::t(4);
            const double Gam = t * (1. - usqr);
            const double eqg0 = t * (Pp - rho3 * usqr);
            // the layered variant drives the rest population with grad lap rho = slope * grad lap phi (twoLayeredFlow2D.h:595-598)
            double uF0 = uF;
            if (mp.sc_force == CLBM_HCZ_FORCE_LAYERED) {
                const double slope = mp.drho * mp.inv_dphi;
                double fx0, fy0;
                hcz2d_force(mp, rho, slope * glx, slope * gly, fx0, fy0);
                uF0 = u0 * fx0 + u1 * fy0;
            }
            const double fg0 = hw * (-uF0 * Gam - uE * (Gam - t));     // (u.(-E)) as the reference writes it
            const double ff0 = hw3 * uG * Gam;
            FOUT(4)[i] = om1 * f[4] + op * Gam + ff0;
            GOUT(4)[i] = om1 * gg[4] + omega * eqg0 + fg0;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // opposite directions share c.u, c.D, c.Eh, c.Gv up to the sign and are collided in pairs
            const int ko = L9f::opp(k);
            const double t = L9f::t(k);
            const bool axis = (L9f::cx(k) != 0) + (L9f::cy(k) != 0) == 1;
            const double cu = cdot<L9f>(k, u0, u1, 0.0);
            const double cD = cdot<L9f>(k, D0, D1, 0.0);
            const double cE = cdot<L9f>(k, E0, E1, 0.0);
            const double cG = cdot<L9f>(k, G0, G1, 0.0);
            const double ev = fma(4.5 * cu, cu, -usqr);
            const double Ge = fma(t, ev, t), Go = (3. * t) * cu;
            const double Gp = Ge + Go, Gm = Ge - Go;
            const double we = fma(axis ? Ba : Bd, ev, axis ? Aa : Ad), wo = orho3 * Go;
            const double pfp = fma(Gp, opg + cG, om1 * f[k]);
            const double pfm = fma(Gm, opg - cG, om1 * f[ko]);
            const double pgp = fma(t, cE - uEh, fma(Gp, cD - uD, fma(om1, gg[k], we + wo)));
            const double pgm = fma(t, -cE - uEh, fma(Gm, -cD - uD, fma(om1, gg[ko], we - wo)));
            push(k, pfp, pgp);
            push(ko, pfm, pgm);
        }
        };
        if (walls_near(x)) collide(std::true_type{}); else collide(std::false_type{});
    }
    if (MULTI && step + 1 < nsteps) cooperative_groups::this_grid().sync();
  }
}

template <int NT, int MINB, bool MRT = false>
static int launch_hcz2d_fused(clbm_ctx *c, int x_begin, int x_end)
{
    const Geom &g = c->geo;
    const int ncol = x_end - x_begin;
    if (ncol <= 0) return 0;
    const int segs = (g.ny + (NT - 4) - 1) / (NT - 4);
    // short x-chunks keep concurrently resident CTAs on neighbouring columns (L2 locality of the overlapping segment
    // rows): 48-64 columns measured best at 2048 x 8194 (14.1 vs 12.5 GLUPS at 128 and 9.6 at 512)
    int xchunk = ncol < 48 ? ncol : 48;
    // small lattices (BASELINE configs[1], 256 x 1026) are L2 resident and latency bound: shorter chunks until there are two
    // CTAs per SM slot, down to 4 columns (tools/small_lattice_chunks.py at 256 x 1026, bit-identical populations: 29.9 us per
    // step at 8 columns, 24.5 at 4, 26.4 at 2, 31.9 at 1 -- the 4-column prologue of this kernel costs more than the SC one)
    const long long want = 2LL * 148 * MINB;
    while (xchunk > 4 && (long long)segs * ((ncol + xchunk - 1) / xchunk) < want) xchunk = xchunk / 2 > 4 ? xchunk / 2 : 4;
    if (c->env.hcz2d_xchunk > 0) xchunk = c->env.hcz2d_xchunk < ncol ? c->env.hcz2d_xchunk : ncol;
    dim3 grid(segs, (ncol + xchunk - 1) / xchunk);
    Hcz2dTables P;
    for (int k = 0; k < 9; ++k) {
        P.fin[k] = c->pop[0][c->parity] + (size_t)k * g.ncs;
        P.gin[k] = c->pop[1][c->parity] + (size_t)k * g.ncs;
        P.fout[k] = c->pop[0][1 - c->parity] + (size_t)k * g.ncs;
        P.gout[k] = c->pop[1][1 - c->parity] + (size_t)k * g.ncs;
    }
    LaunchScope ls(c, "hcz2d_fused_collide_stream", ncol * 2 >= g.nx);   // the boundary-column launches of the overlap protocol are not the dominant kernel
    hcz2d_fused_kernel<NT, MINB, MRT><<<grid, NT, 0, c->stream>>>(P, c->flag, c->fld[0], g, c->mp, xchunk, x_begin, x_end, 1, nullptr);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

bool hcz2d_fused_eligible(const clbm_ctx *c) { return c->geo.ny >= 4 && c->geo.ncs < (1LL << 31); }

// nsteps steps in one cooperative launch (MULTI form); *done stays 0 when the lattice does not qualify: the grid must be
// co-resident (grid barrier) and both population buffers should live in L2 -- at HBM size the launch-per-step kernel with its
// 48-column chunks is the faster one (no phi pass, no barriers)
template <int NT, int MINB, bool MRT>
static int launch_hcz2d_multi(clbm_ctx *c, int nsteps, int *done)
{
    const Geom &g = c->geo;
    auto kern = hcz2d_fused_kernel<NT, MINB, MRT, true>;
    int per_sm = 0, sms = 0;
    const size_t smem = 9 * (size_t)NT * sizeof(double);
    CLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CLBM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    CLBM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    const int segs = (g.ny + (NT - 4) - 1) / (NT - 4);
    const int max_chunks = per_sm * sms / segs;
    if (max_chunks < 1) return 0;
    int xchunk = (g.nx + max_chunks - 1) / max_chunks;
    if (xchunk < 2) xchunk = 2;
    if (c->env.hcz2d_xchunk > 0 && c->env.hcz2d_xchunk >= xchunk) xchunk = c->env.hcz2d_xchunk < g.nx ? c->env.hcz2d_xchunk : g.nx;
    dim3 grid(segs, (g.nx + xchunk - 1) / xchunk);
    Hcz2dTables P;
    for (int k = 0; k < 9; ++k) {
        P.fin[k] = c->pop[0][c->parity] + (size_t)k * g.ncs;
        P.gin[k] = c->pop[1][c->parity] + (size_t)k * g.ncs;
        P.fout[k] = c->pop[0][1 - c->parity] + (size_t)k * g.ncs;
        P.gout[k] = c->pop[1][1 - c->parity] + (size_t)k * g.ncs;
    }
    const uint8_t *fl = c->flag;
    const double *phi_g = c->fld[0];
    double *phi_w = c->fld[0];
    Geom gg = g;
    ModelParams mp = c->mp;
    int x_begin = 0, x_end = g.nx;
    void *args[] = {&P, &fl, &phi_g, &gg, &mp, &xchunk, &x_begin, &x_end, &nsteps, &phi_w};
    LaunchScope ls(c, "hcz2d_fused_multi_step", true);
    CLBM_CUDA(cudaLaunchCooperativeKernel((const void *)kern, grid, dim3(NT), args, smem, c->stream));
    if (nsteps & 1) c->parity = 1 - c->parity;
    *done = 1;
    return 0;
}

// HCZ D2Q9 on a single slab: all the steps of a clbm_step(n >= 2) call in one launch, only with CLBM_HCZ2D_MULTI=1 (measured slower
// than the launch-per-step path at BASELINE configs[1], see the kernel's header)
int hcz2d_fused_multi_step(clbm_ctx *c, int nsteps, int *done)
{
    *done = 0;
    if (nsteps < 2 || c->multi || c->prm.fused != 1 || c->env.hcz2d_tile >= 0 || !hcz2d_fused_eligible(c) || c->env.hcz2d_multi != 1) return 0;
    if (c->prm.collision == CLBM_COLLISION_MRT) return launch_hcz2d_multi<128, 3, true>(c, nsteps, done);
    return launch_hcz2d_multi<128, 4, false>(c, nsteps, done);
}

// collide + push of the columns [x_begin, x_end) of this slab (the whole slab, or a range of the overlap protocol)
int hcz2d_fused_range(clbm_ctx *c, int x_begin, int x_end)
{
    int variant = c->prm.fused > 1 ? c->prm.fused : 0;
    if (c->env.hcz2d_tile >= 0) variant = c->env.hcz2d_tile;
    if (c->prm.collision == CLBM_COLLISION_MRT) return launch_hcz2d_fused<128, 3, true>(c, x_begin, x_end);   // 168 registers: no spills
    switch (variant) {
    case 2: return launch_hcz2d_fused<64, 8>(c, x_begin, x_end);
    case 3: return launch_hcz2d_fused<96, 4>(c, x_begin, x_end);
    case 4: return launch_hcz2d_fused<128, 3>(c, x_begin, x_end);
    default: return launch_hcz2d_fused<128, 4>(c, x_begin, x_end);   // 128 registers, 16 warps per SM: best of the sweep at 2048 x 8194 with the cp.async staging
    }
}

int hcz2d_fused_launch(clbm_ctx *c) { return hcz2d_fused_range(c, 0, c->geo.nx); }

}  // namespace clbm
