// sc_fused.cu -- fused plane-marching Shan-Chen time step: ONE kernel per step, every population
// read once and written once from HBM.
//
// The staged form (sc_kernels.cu) reads the populations twice: once to build psi(rho), once to
// collide.  Here a thread block owns a TY x TZ tile of the (y,z) plane and marches along x (the
// slowest index).  While plane x is collided, the populations of plane x+1 are already in
// registers: they were loaded to compute psi(x+1) into a 4-slot shared-memory ring of
// (TY+2) x (TZ+2) psi planes with halos, and they are kept for the collision of the next
// iteration.  Only the halo ring of each plane is read a second time (by the neighbouring
// tile), which the L2 absorbs.  One __syncthreads per plane.
//
//   psi ring value < 0  <=>  bounce_back node (psi >= 0 everywhere else), so the ring also
//   carries the node mask for the wall term and for the half-way bounce-back of the push.
//
// Physics per cell: sc_cell.cuh (force, tau-shifted BGK; SC/apps/laplace2D.h:198-306,
// SC/apps/contactAngle2D.h:248-355).
#include <cooperative_groups.h>

#include <cstdlib>

#include "sc_cell.cuh"

namespace clbm {

template <class L, int TY, int TZ>
struct FusedCfg {
    static constexpr int NT = TY * TZ;
    static constexpr int SY = TY + 2;
    static constexpr int SZ = (L::D == 3) ? TZ + 2 : 1;
    static constexpr int NHALO = (L::D == 3) ? 2 * SZ + 2 * TY : 2;
    static_assert(L::D == 3 || TZ == 1, "D2Q9 tiles are one-dimensional");
};

template <class L> struct PopTable {
    const double *in[L::Q];   // fin + k*ncs  (constant-bank operands: one IMAD.WIDE per address)
    double *out[L::Q];        // fout + k*ncs
};

// GUO = true: the Rayleigh-Taylor variant (SC/apps/RayleighTaylor2D.h; psi = 1 - exp(-rho), a wall neighbour
// contributes the psi of the opposite neighbour, Guo forcing) -- a compile-time flag, the Yuan-CS code is unchanged.
// MRT = true: CLBM_COLLISION_MRT for D2Q9 (sc_collide_mrt), likewise a compile-time flag.
//
// MULTI = true: the whole grid is resident (cooperative launch) and runs `nsteps` time steps in ONE launch, a grid barrier and a
// swap of the two population buffers between them.  For lattices that live in L2 (BASELINE configs[0]: 256 x 256, 4.7 MB per
// buffer) a step is a few microseconds of latency and the kernel boundary is most of it: 8.2 us per step launch by launch.
template <class L, int TY, int TZ, int MINB, bool GUO = false, bool MRT = false, int PF = 2, bool MULTI = false>
__global__ void __launch_bounds__(TY *TZ, MINB)
sc_fused_kernel(const PopTable<L> P0, const uint8_t *__restrict__ flag, const double *__restrict__ psi_g, Geom g,
                ModelParams mp, int xchunk, int nsteps)
{
  for (int step = 0; step < (MULTI ? nsteps : 1); ++step) {
    // population tables of this step: the buffers swap roles every step (the stores of the previous step are visible behind the
    // grid barrier; the loads below are plain loads, never the non-coherent path, because `in` was written by this very kernel)
    PopTable<L> P;
    if (MULTI && (step & 1)) {
#pragma unroll
        for (int k = 0; k < L::Q; ++k) { P.in[k] = P0.out[k]; P.out[k] = const_cast<double *>(P0.in[k]); }
    } else {
#pragma unroll
        for (int k = 0; k < L::Q; ++k) { P.in[k] = P0.in[k]; P.out[k] = P0.out[k]; }
    }
    using C = FusedCfg<L, TY, TZ>;
    __shared__ double psi_s[4][C::SY][C::SZ];

    const int tid = threadIdx.x;
    const int tz = tid % TZ, ty = tid / TZ;
    const int y0 = blockIdx.y * TY, z0 = blockIdx.x * TZ;
    const int y = y0 + ty, z = z0 + tz;
    const bool inside = (y < g.ny) && (z < g.nz);
    const int xa = blockIdx.z * xchunk;
    const int xb = min(g.nx, xa + xchunk);
    const int plane = (int)g.plane, nz = g.nz, G = g.G;
    // rows / columns of this tile that lie inside the lattice (partial tiles at the upper edges);
    // the halo ring hugs them: rows sy = 0 and ty_n+1, columns sz = 0 and tz_n+1
    const int ty_n = min(TY, g.ny - y0), tz_n = (L::D == 3) ? min(TZ, g.nz - z0) : 1;
    const int nrow = tz_n + 2;
    const int nhalo = (L::D == 3) ? 2 * nrow + 2 * ty_n : 2;
    const int cz0 = (L::D == 3) ? tz + 1 : 0;
    const int yz = y * nz + z;   // in-plane index of the own cell

    // psi (or -1 for a wall) of storage plane xs into ring slot `slot`.  The thread's own populations stay in
    // fk, its own psi / G1 branch in ps / gp (used when that plane is collided one iteration later).
    auto fill = [&](int xs, int slot, double *fk, double &ps, bool &gp) {
        if (!g.wrapx && (xs < G || xs >= g.nx + G)) {
            // x-slab mode: ghost plane of the neighbour slab -> psi from the exchanged moment halo, mask from flag[]
            if (inside) {
                const int i = xs * plane + yz;
                psi_s[slot][ty + 1][cz0] = (flag[i] == CELL_BB) ? -1.0 : fabs(psi_g[i]);
            }
            for (int h = tid; h < nhalo; h += C::NT) {
                int sy, sz;
                if (L::D == 3) {
                    if (h < nrow) { sy = 0; sz = h; }
                    else if (h < 2 * nrow) { sy = ty_n + 1; sz = h - nrow; }
                    else { const int q = h - 2 * nrow; sy = 1 + (q >> 1); sz = (q & 1) ? tz_n + 1 : 0; }
                } else { sy = h ? ty_n + 1 : 0; sz = 0; }
                const int yy = g.wy(y0 + sy - 1), zz = (L::D == 3) ? g.wz(z0 + sz - 1) : 0;
                const int i = xs * plane + yy * nz + zz;
                psi_s[slot][sy][sz] = (flag[i] == CELL_BB) ? -1.0 : fabs(psi_g[i]);
            }
            return;
        }
        if (inside) {
            const int i = xs * plane + yz;
#pragma unroll
            for (int k = 0; k < L::Q; ++k) fk[k] = MULTI ? __ldcg(P.in[k] + i) : P.in[k][i];
            double v = -1.0;
            ps = 0.0;
            gp = true;
            if (flag[i] != CELL_BB) {
                if constexpr (GUO) ps = scrt_psi(Mom<L>::sum(fk));
                else ps = sc_psi_g1(mp, Mom<L>::sum(fk), gp);
                v = ps;
            }
            psi_s[slot][ty + 1][cz0] = v;
        }
        for (int h = tid; h < nhalo; h += C::NT) {
            int sy, sz;
            if (L::D == 3) {
                if (h < nrow) { sy = 0; sz = h; }
                else if (h < 2 * nrow) { sy = ty_n + 1; sz = h - nrow; }
                else { const int q = h - 2 * nrow; sy = 1 + (q >> 1); sz = (q & 1) ? tz_n + 1 : 0; }
            } else { sy = h ? ty_n + 1 : 0; sz = 0; }
            const int yy = g.wy(y0 + sy - 1), zz = (L::D == 3) ? g.wz(z0 + sz - 1) : 0;
            const int i = xs * plane + yy * nz + zz;
            double v = -1.0;
            // populations fetched whatever the mask says (a bounce_back cell's are valid memory, just unused): with the loads behind
            // the mask test a halo thread paid two memory latencies in a row while its whole CTA waited at the barrier
            const uint8_t flh = flag[i];
            double fh[L::Q];
#pragma unroll
            for (int k = 0; k < L::Q; ++k) fh[k] = MULTI ? __ldcg(P.in[k] + i) : P.in[k][i];
            if (flh != CELL_BB) {
                bool gph;
                if constexpr (GUO) v = scrt_psi(Mom<L>::sum(fh));
                else v = sc_psi_g1(mp, Mom<L>::sum(fh), gph);
            }
            psi_s[slot][sy][sz] = v;
        }
    };

    double fc[L::Q], fn[L::Q];
    double psc, psn;
    bool gpc, gpn;
    fill(g.wx(xa - 1) + G, (xa + 3) & 3, fn, psn, gpn);   // plane xa-1: only its psi ring is needed
    fill(xa + G, xa & 3, fc, psc, gpc);

    // in-plane neighbour offsets of this thread's column (x offsets change per plane)
    const int oym = (g.wy(y - 1) - y) * nz, oyp = (g.wy(y + 1) - y) * nz;
    const int ozm = g.wz(z - 1) - z, ozp = g.wz(z + 1) - z;

    for (int x = xa; x < xb; ++x) {
        const int xp = g.wx(x + 1), xm = g.wx(x - 1);
        fill(xp + G, (x + 1) & 3, fn, psn, gpn);
        // The loads of fill() are consumed at once (psi needs the density), so every plane pays a full memory latency with
        // only 16 warps per SM to hide it (ncu at 8192^2, D2Q9: DRAM at 47 %, top stall long_scoreboard).  Pull the plane
        // two further on into L2 now: those loads then wait for an L2 hit instead of HBM.  No registers, no shared memory.
        if (PF > 0 && inside) {
            const int xq = g.wrapx ? g.wx(g.wx(x + 1) + PF) : x + 1 + PF;
            if (xq + G < g.nx + 2 * G) {
                const int i = (xq + G) * plane + yz;
#pragma unroll
                for (int k = 0; k < L::Q; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.in[k] + i));
            }
        }
        __syncthreads();

        const int sm = (x + 3) & 3, s0 = x & 3, sp = (x + 1) & 3;
        if (inside && psi_s[s0][ty + 1][cz0] >= 0.0) {
            ScForceSums s = {{0., 0., 0.}, {0., 0., 0.}, 0u};
#pragma unroll
            for (int k = 0; k < L::Q; ++k) {
                if (k == L::REST) continue;
                const int slot = L::cx(k) < 0 ? sm : (L::cx(k) > 0 ? sp : s0);
                const double v = psi_s[slot][ty + 1 + L::cy(k)][cz0 + L::cz(k)];
                if constexpr (GUO) {
                    // wall neighbour: psi of the opposite neighbour (RayleighTaylor2D.h:246-262); a wall there too holds psi = 0
                    const int oslot = L::cx(k) < 0 ? sp : (L::cx(k) > 0 ? sm : s0);
                    const double vo = psi_s[oslot][ty + 1 - L::cy(k)][cz0 - L::cz(k)];
                    if (v < 0.0) s.wall |= 1u << k;
                    sc_force_add<L>(s, k, false, v < 0.0 ? fmax(vo, 0.0) : v);
                } else {
                    sc_force_add<L>(s, k, v < 0.0, v);
                }
            }
            double out[L::Q];
            if constexpr (GUO && MRT) scrt_collide_mrt<L>(mp, fc, s, Mom<L>::sum(fc), psc, out);
            else if constexpr (GUO) scrt_collide<L>(mp, fc, s, Mom<L>::sum(fc), psc, out);
            else if constexpr (MRT) sc_collide_mrt<L>(mp, fc, s, Mom<L>::sum(fc), psc, gpc, out);
            else sc_collide<L>(mp, fc, s, psc, gpc, out);

            const int i = (x + G) * plane + yz;
            const int oxm = (xm - x) * plane, oxp = (xp - x) * plane;
#pragma unroll
            for (int k = 0; k < L::Q; ++k) {
                if (k == L::REST) { P.out[k][i] = out[k]; continue; }
                const int off = (L::cx(k) < 0 ? oxm : (L::cx(k) > 0 ? oxp : 0)) + (L::cy(k) < 0 ? oym : (L::cy(k) > 0 ? oyp : 0)) +
                                (L::cz(k) < 0 ? ozm : (L::cz(k) > 0 ? ozp : 0));
                if (s.wall & (1u << k)) P.out[L::opp(k)][i] = out[k];   // half-way bounce-back
                else P.out[k][i + off] = out[k];
            }
        }
#pragma unroll
        for (int k = 0; k < L::Q; ++k) fc[k] = fn[k];
        psc = psn;
        gpc = gpn;
    }
    if (MULTI && step + 1 < nsteps) cooperative_groups::this_grid().sync();
  }
}

struct FusedChoice { int ty, tz; };

template <class L, int TY, int TZ, int MINB, bool GUO = false, bool MRT = false>
static int launch_fused(clbm_ctx *c)
{
    const Geom &g = c->geo;
    const int tiles = ((g.ny + TY - 1) / TY) * ((g.nz + TZ - 1) / TZ);
    // short x-chunks: concurrently resident CTAs stay on neighbouring planes and share their halo rows through L2
    // (32 planes measured best for D2Q9 at 8192^2: 25.2 vs 22.0 GLUPS with 512-plane chunks); shorter still when that
    // is needed to fill the SMs
    int xchunk = g.nx < 32 ? g.nx : 32;
    const long long want = 2LL * 148 * MINB;
    if ((long long)tiles * ((g.nx + xchunk - 1) / xchunk) < want) {
        const long long nch = (want + tiles - 1) / tiles;
        xchunk = (int)((g.nx + nch - 1) / nch);
        // Lattices this small are L2 resident and LATENCY bound: a CTA marches its chunk serially, so shorter chunks mean
        // more CTAs in flight; the extra prologue planes are L2 hits.  D2Q9 single slab, measured at 256 x 256
        // (tools/small_lattice_chunks.py, bit-identical populations): 18.4 us per step at 8 columns, 12.3 at 4, 10.1 at 2,
        // 8.2 at 1 (the staged kernels: 10.2).  Other paths keep the 8-plane floor they were measured with.
        const int floor_ = (L::D == 2 && !c->multi) ? 1 : 8;
        if (xchunk < floor_) xchunk = g.nx < floor_ ? g.nx : floor_;
    }
    if (c->env.sc_xchunk > 0) xchunk = c->env.sc_xchunk < g.nx ? c->env.sc_xchunk : g.nx;
    dim3 grid((g.nz + TZ - 1) / TZ, (g.ny + TY - 1) / TY, (g.nx + xchunk - 1) / xchunk);
    PopTable<L> P;
    for (int k = 0; k < L::Q; ++k) {
        P.in[k] = c->pop[0][c->parity] + (size_t)k * g.ncs;
        P.out[k] = c->pop[0][1 - c->parity] + (size_t)k * g.ncs;
    }
    LaunchScope ls(c, "sc_fused_collide_stream", true);
    sc_fused_kernel<L, TY, TZ, MINB, GUO, MRT><<<grid, TY * TZ, 0, c->stream>>>(P, c->flag, c->fld[0], g, c->mp, xchunk, 1);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

// nsteps steps of a small D2Q9 lattice in one cooperative launch (MULTI form of the kernel); *done = 0 when the lattice does not
// qualify (the caller then steps launch by launch).  The parity advances by nsteps.
template <int TY, int MINB>
static int launch_fused_multi(clbm_ctx *c, int nsteps, int xchunk, int *done)
{
    const Geom &g = c->geo;
    auto kern = sc_fused_kernel<D2Q9, TY, 1, MINB, false, false, 0, true>;
    int per_sm = 0;
    CLBM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TY, 0));
    int sms = 0;
    CLBM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    dim3 grid(1, (g.ny + TY - 1) / TY, (g.nx + xchunk - 1) / xchunk);
    if ((long long)grid.y * grid.z > (long long)per_sm * sms) return 0;   // not co-resident: no grid barrier
    PopTable<D2Q9> P;
    for (int k = 0; k < 9; ++k) {
        P.in[k] = c->pop[0][c->parity] + (size_t)k * g.ncs;
        P.out[k] = c->pop[0][1 - c->parity] + (size_t)k * g.ncs;
    }
    const uint8_t *fl = c->flag;
    const double *psi = c->fld[0];
    Geom gg = g;
    ModelParams mp = c->mp;
    void *args[] = {&P, &fl, &psi, &gg, &mp, &xchunk, &nsteps};
    LaunchScope ls(c, "sc_fused_multi_step", true);
    CLBM_CUDA(cudaLaunchCooperativeKernel((const void *)kern, grid, dim3(TY, 1, 1), args, 0, c->stream));
    if (nsteps & 1) c->parity = 1 - c->parity;
    *done = 1;
    return 0;
}

int sc2d_resident_multi_step(clbm_ctx *c, int nsteps, int *done);

// Yuan-CS BGK Shan-Chen D2Q9 on a single slab that fits in L2: several steps per launch (CLBM_SC_MULTI=0 turns it off)
int sc_fused_multi_step(clbm_ctx *c, int nsteps, int *done)
{
    *done = 0;
    const Geom &g = c->geo;
    if (c->Q != 9 || c->multi || !c->prm.fused || c->prm.fused > 1 || c->env.sc_tile >= 0 || c->prm.collision != CLBM_COLLISION_BGK ||
        c->mp.sc_force == CLBM_SC_FORCE_EXPGUO || c->profiling || c->ktiming || nsteps < 2)
        return 0;
    if (c->env.sc_multi == 0) return 0;
    if ((size_t)g.ncs * 9 * sizeof(double) * 2 > 64u << 20) return 0;   // both buffers well inside the L2
    if (c->env.sc_multi < 0 || c->env.sc_multi >= 4) {   // default: the column-resident kernel below, where the lattice qualifies
        if (int rc = sc2d_resident_multi_step(c, nsteps, done)) return rc;
        if (*done) return 0;
    }
    if (c->env.sc_multi == 1 || c->env.sc_multi == 3) {   // tile-height experiments of the plane-marching form
        const int xchunk = c->env.sc_xchunk > 0 ? (c->env.sc_xchunk < g.nx ? c->env.sc_xchunk : g.nx) : 1;
        if (c->env.sc_multi == 3) return launch_fused_multi<64, 8>(c, nsteps, xchunk, done);
        return launch_fused_multi<128, 4>(c, nsteps, xchunk, done);
    }
    // too many columns to keep one CTA per column resident: the plane-marching form with the shortest x-chunks that still fit
    // (512 x 256: 7.1 us per step at 2 columns per CTA against 14.3 launch by launch)
    for (int xchunk = c->env.sc_xchunk > 0 ? c->env.sc_xchunk : 1; xchunk <= 8 && xchunk <= g.nx; xchunk *= 2) {
        if (int rc = launch_fused_multi<256, 2>(c, nsteps, xchunk, done)) return rc;
        if (*done || c->env.sc_xchunk > 0) return 0;
    }
    return 0;
}

bool sc_tma_eligible(const clbm_ctx *c);            // sc_fused_tma.cu
int sc_fused_tma_step(clbm_ctx *c, int variant);
int sc_fused_tma_range(clbm_ctx *c, int variant, int x_begin, int x_end, int x2_begin, int x2_end);
int sc_fused_tma_persist_range(clbm_ctx *c, int variant, int x_begin, int x_end, int x2_begin, int x2_end);   // sc_fused_tma_persist.cu
bool sc2d_tma_eligible(const clbm_ctx *c);          // sc2d_tma.cu
int sc2d_tma_step(clbm_ctx *c, int variant);

// one fused collide-stream sweep over the local planes; does NOT flip the parity
// tile variant of the TMA kernel this context would run, 0 when it runs one of the register-pipelined kernels
static int sc_tma_variant(const clbm_ctx *c)
{
    // clbm_params.fused: 1 = default fused kernel, >1 = explicit tile variant (tuning / tests); env overrides
    if (c->prm.collision == CLBM_COLLISION_MRT) return 0;   // the MRT operator lives in the register-pipelined and the staged kernels
    int variant = c->prm.fused > 1 ? c->prm.fused : 0;
    if (c->env.sc_tile >= 0) variant = c->env.sc_tile;
    // D3Q19 default: the TMA-staged kernel with 8 x 64 tiles as two independent 4-row groups, early stage release (variant 29, = 24 with 32-bit store indices: best of
    // the sweeps in profiles/README.md)
    if (variant == 0 && c->Q == 19 && sc_tma_eligible(c)) variant = 29;
    return (variant >= 10 && sc_tma_eligible(c)) ? variant : 0;
}

// the work queue of the next launch of the persistent kernel (allocated and zeroed on first use)
int sc_persist_queue(clbm_ctx *c, int **q)
{
    if (!c->sc_queue) {
        CLBM_CUDA(cudaMalloc(&c->sc_queue, 16 * sizeof(int)));
        CLBM_CUDA(cudaMemsetAsync(c->sc_queue, 0, 16 * sizeof(int), c->stream));
        if (c->stream_b) {   // the boundary stream may launch the kernel too: order the memset before anything it does
            CLBM_CUDA(cudaStreamSynchronize(c->stream));
        }
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, c->device) == cudaSuccess) c->sm_count = n;
    }
    *q = c->sc_queue + 2 * (c->sc_queue_next++ & 7);
    return 0;
}

// x-range launches (overlap protocol of the slab exchange) exist for the TMA kernel
bool sc_range_supported(const clbm_ctx *c) { return c->prm.fused && sc_tma_variant(c) != 0; }
int sc_collide_range_fused(clbm_ctx *c, int x_begin, int x_end, int x2_begin, int x2_end)
{
    const int v = sc_tma_variant(c);
    if (v >= 40) return sc_fused_tma_persist_range(c, v, x_begin, x_end, x2_begin, x2_end);
    return sc_fused_tma_range(c, v, x_begin, x_end, x2_begin, x2_end);
}

int sc_fused_launch(clbm_ctx *c)
{
    int rc;
    const bool mrt = c->prm.collision == CLBM_COLLISION_MRT;
    int variant = c->prm.fused > 1 ? c->prm.fused : 0;
    if (c->env.sc_tile >= 0) variant = c->env.sc_tile;
    if (mrt) variant = 0;
    if (variant == 0 && c->Q == 19 && sc_tma_eligible(c) && !mrt) variant = 29;
    if (variant >= 40 && sc_tma_eligible(c)) return sc_fused_tma_persist_range(c, variant, 0, c->geo.nx, 0, 0);
    if (variant >= 10 && sc_tma_eligible(c)) return sc_fused_tma_step(c, variant);
    if (variant >= 10) variant = 0;
    // D2Q9 at HBM size: the TMA-staged column kernel (sc2d_tma.cu; bit-identical populations).  L2-resident lattices keep the
    // register-pipelined kernel with its one-column chunks (and clbm_step(n >= 2) the multi-step launches below).
    if (c->Q == 9 && variant == 0 && !mrt && c->env.sc2d_tma != 0 && sc2d_tma_eligible(c) &&
        (c->env.sc2d_tma > 0 || (size_t)c->geo.ncs * 9 * sizeof(double) * 2 > (size_t)256 << 20))
        return sc2d_tma_step(c, c->env.sc2d_tma > 1 ? c->env.sc2d_tma : 0);
    if (c->mp.sc_force == CLBM_SC_FORCE_EXPGUO && mrt) return launch_fused<D2Q9, 128, 1, 3, true, true>(c);
    if (c->mp.sc_force == CLBM_SC_FORCE_EXPGUO) return launch_fused<D2Q9, 128, 1, 4, true>(c);   // D2Q9 only (clbm_create)
    if (mrt && c->Q == 9) return launch_fused<D2Q9, 128, 1, 3, false, true>(c);
    if (mrt) return launch_fused<D3Q19, 4, 64, 1, false, true>(c);   // D3Q19 MRT: one CTA per SM, up to 255 registers for the 19 x 19 transforms
    if (c->Q == 9) {
        switch (variant) {
        case 1: rc = launch_fused<D2Q9, 256, 1, 2>(c); break;
        case 2: rc = launch_fused<D2Q9, 64, 1, 8>(c); break;
        default: rc = launch_fused<D2Q9, 128, 1, 4>(c); break;
        }
    } else {
        switch (variant) {
        case 1: rc = launch_fused<D3Q19, 8, 64, 1>(c); break;
        case 2: rc = launch_fused<D3Q19, 8, 32, 2>(c); break;
        case 3: rc = launch_fused<D3Q19, 16, 32, 1>(c); break;
        case 4: rc = launch_fused<D3Q19, 2, 128, 2>(c); break;
        case 5: rc = launch_fused<D3Q19, 6, 64, 1>(c); break;
        case 6: rc = launch_fused<D3Q19, 4, 64, 1>(c); break;
        case 7: rc = launch_fused<D3Q19, 3, 128, 1>(c); break;
        case 8: rc = launch_fused<D3Q19, 4, 128, 1>(c); break;
        default: rc = launch_fused<D3Q19, 4, 64, 2>(c); break;
        }
    }
    return rc;
}

int sc_fused_step(clbm_ctx *c)
{
    int rc = sc_fused_launch(c);
    if (rc) return rc;
    c->parity = 1 - c->parity;
    return 0;
}

}  // namespace clbm

// ---- L2-resident D2Q9 lattices: one CTA per lattice column, one thread per node, every step of a clbm_step(n) call in ONE
// cooperative launch.  BASELINE configs[0] (256 x 256) is 4.7 MB per buffer: it never leaves L2, and a step is a chain of
// latencies -- launch, three dependent load / psi rounds of the plane-marching kernel, the stores.  Here a thread issues the 27
// loads of its node in the three columns x-1, x, x+1 at once (plain L2 loads: the data was written by this kernel one step
// earlier), builds the three psi values, meets its column in shared memory, collides and pushes; a grid barrier ends the
// step.  The node masks of the three columns are read once, before the first step.  Same per-cell functions, same summation
// order as sc_fused_kernel: the populations are bit-identical (tools/small_lattice_multi.py).
namespace clbm {

//
// P2P = true: no grid barrier.  Column x of step t + 1 depends on what the CTAs x-2 .. x+2 did in step t (its three input columns
// are written by x-2 .. x+2, and the columns it overwrites are still being read by them), so every CTA publishes the number of
// steps it has completed and waits for those four neighbours only: one flag round trip through L2 instead of an atomic on a
// word that all CTAs hammer.  The flags keep counting across launches (`epoch` = steps completed before this launch).
template <int NT, bool P2P>
__global__ void __launch_bounds__(NT)
sc2d_resident_kernel(const PopTable<D2Q9> P0, const uint8_t *__restrict__ flag, Geom g, ModelParams mp, int nsteps, int *progress, int epoch)
{
    using L = D2Q9;
    extern __shared__ double psi_col[];          // [3][ny + 2]: psi (or -1 for a wall) of columns x-1, x, x+1 with the periodic rows
    const int ny = g.ny, G = g.G, plane = (int)g.plane;
    const int x = blockIdx.x, y = threadIdx.x;
    const bool act = y < ny;
    const int xm = g.wx(x - 1), xp = g.wx(x + 1);
    const int im = (xm + G) * plane + y, ic = (x + G) * plane + y, ip = (xp + G) * plane + y;
    uint8_t flm = CELL_BB, flc = CELL_BB, flp = CELL_BB;
    if (act) { flm = flag[im]; flc = flag[ic]; flp = flag[ip]; }
    double *pm = psi_col, *pc = psi_col + (ny + 2), *pp = psi_col + 2 * (ny + 2);
    const int oym = g.wy(y - 1) - y, oyp = g.wy(y + 1) - y;
    const int oxm = (xm - x) * plane, oxp = (xp - x) * plane;
    auto put = [&](double *col, double v) {
        col[y + 1] = v;
        if (y == ny - 1) col[0] = v;        // periodic rows (walls are mask rows: a bulk node never reads across them)
        if (y == 0) col[ny + 1] = v;
    };

    for (int step = 0; step < nsteps; ++step) {
        const bool odd = step & 1;
        double fc[9];
        double psc = 0.0;
        bool gpc = true;
        if (act) {
            double fm[9], fp[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const double *src = odd ? P0.out[k] : P0.in[k];
                fm[k] = __ldcg(src + im);
                fc[k] = __ldcg(src + ic);
                fp[k] = __ldcg(src + ip);
            }
            bool gpx;
            put(pm, flm == CELL_BB ? -1.0 : sc_psi_g1(mp, Mom<L>::sum(fm), gpx));
            put(pp, flp == CELL_BB ? -1.0 : sc_psi_g1(mp, Mom<L>::sum(fp), gpx));
            if (flc != CELL_BB) psc = sc_psi_g1(mp, Mom<L>::sum(fc), gpc);
            put(pc, flc == CELL_BB ? -1.0 : psc);
        }
        __syncthreads();
        {
            if (act && flc != CELL_BB) {
                ScForceSums s = {{0., 0., 0.}, {0., 0., 0.}, 0u};
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    if (k == L::REST) continue;
                    const double *col = L::cx(k) < 0 ? pm : (L::cx(k) > 0 ? pp : pc);
                    const double v = col[y + 1 + L::cy(k)];
                    sc_force_add<L>(s, k, v < 0.0, v);
                }
                double out[9];
                sc_collide<L>(mp, fc, s, psc, gpc, out);
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    double *dst = odd ? const_cast<double *>(P0.in[k]) : P0.out[k];
                    double *dopp = odd ? const_cast<double *>(P0.in[L::opp(k)]) : P0.out[L::opp(k)];
                    if (k == L::REST) { dst[ic] = out[k]; continue; }
                    const int off = (L::cx(k) < 0 ? oxm : (L::cx(k) > 0 ? oxp : 0)) + (L::cy(k) < 0 ? oym : (L::cy(k) > 0 ? oyp : 0));
                    if (s.wall & (1u << k)) dopp[ic] = out[k];   // half-way bounce-back
                    else dst[ic + off] = out[k];
                }
            }
        }
        if (P2P) {
            __syncthreads();                                    // every store of this CTA's step is issued
            if (y == 0) {
                __threadfence();
                *(volatile int *)&progress[x] = epoch + step + 1;
            }
            if (step + 1 < nsteps) {
                if (y < 4) {
                    const int d = y < 2 ? y - 2 : y - 1;          // -2, -1, +1, +2
                    int xn = x + d;
                    xn = xn < 0 ? xn + g.nx : (xn >= g.nx ? xn - g.nx : xn);
                    const volatile int *pf = &progress[xn];
                    while (*pf - (epoch + step + 1) < 0) { }
                    __threadfence();
                }
                __syncthreads();
            }
        } else if (step + 1 < nsteps) {
            cooperative_groups::this_grid().sync();
        }
    }
}

template <int NT, bool P2P>
static int launch_resident(clbm_ctx *c, int nsteps, int *done)
{
    const Geom &g = c->geo;
    auto kern = sc2d_resident_kernel<NT, P2P>;
    if (P2P && g.nx < 5) return 0;
    if (P2P && !c->resident_progress) {
        CLBM_CUDA(cudaMalloc(&c->resident_progress, (size_t)g.nx * sizeof(int)));
        CLBM_CUDA(cudaMemsetAsync(c->resident_progress, 0, (size_t)g.nx * sizeof(int), c->stream));
        c->resident_epoch = 0;
    }
    const size_t smem = 3 * (size_t)(g.ny + 2) * sizeof(double);
    int per_sm = 0, sms = 0;
    CLBM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    CLBM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    if ((long long)g.nx > (long long)per_sm * sms) return 0;   // not co-resident: no grid barrier
    PopTable<D2Q9> P;
    for (int k = 0; k < 9; ++k) {
        P.in[k] = c->pop[0][c->parity] + (size_t)k * g.ncs;
        P.out[k] = c->pop[0][1 - c->parity] + (size_t)k * g.ncs;
    }
    const uint8_t *fl = c->flag;
    Geom gg = g;
    ModelParams mp = c->mp;
    int *prog = c->resident_progress;
    int epoch = c->resident_epoch;
    void *args[] = {&P, &fl, &gg, &mp, &nsteps, &prog, &epoch};
    LaunchScope ls(c, "sc2d_resident_multi_step", true);
    CLBM_CUDA(cudaLaunchCooperativeKernel((const void *)kern, dim3(g.nx), dim3(NT), args, smem, c->stream));
    if (P2P) c->resident_epoch += nsteps;
    if (nsteps & 1) c->parity = 1 - c->parity;
    *done = 1;
    return 0;
}

int sc2d_resident_multi_step(clbm_ctx *c, int nsteps, int *done)
{
    const int ny = c->geo.ny;
    if (c->env.sc_multi == 6) {   // neighbour flags instead of the grid barrier: measured SLOWER (8.5 against 4.5 us per step at 256 x 256:
                                  // every step waits for the slowest of four neighbours, and a flag is seen one poll round trip late)
        if (ny <= 128) return launch_resident<128, true>(c, nsteps, done);
        if (ny <= 256) return launch_resident<256, true>(c, nsteps, done);
        if (ny <= 512) return launch_resident<512, true>(c, nsteps, done);
        if (ny <= 1024) return launch_resident<1024, true>(c, nsteps, done);
        return 0;
    }
    if (ny <= 128) return launch_resident<128, false>(c, nsteps, done);
    if (ny <= 256) return launch_resident<256, false>(c, nsteps, done);
    if (ny <= 512) return launch_resident<512, false>(c, nsteps, done);
    if (ny <= 1024) return launch_resident<1024, false>(c, nsteps, done);
    return 0;
}

}  // namespace clbm
