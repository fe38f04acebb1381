// sc_fused_tma.cu -- D3Q19 Shan-Chen fused step with TMA-staged population tiles (sm_100a).
//
// Same plane-marching scheme as sc_fused.cu (one launch per time step, every population read once and
// written once from HBM), but the loads are taken off the instruction stream:
//
//   * one cp.async.bulk.tensor (TMA) per x-plane fetches the 19 x (TY+2) x (TZ+2) box "tile + halo ring,
//     all directions" of the population tensor [k][x][y][z] into shared memory; completion is signalled
//     on an mbarrier (expect_tx / complete_tx).  Two stages: the box of plane x+2 is in flight while
//     plane x is collided, so HBM latency is hidden without holding a second population set in registers.
//   * psi(rho) of plane x+1 (tile + halo) is computed straight from the staged box into a 4-slot ring of
//     psi planes; the force stencil of plane x reads ring slots x-1, x, x+1.
//   * the collision takes the 19 own populations of plane x from the stage (LDS), the post-collision
//     values are pushed to the neighbours with plain coalesced stores (half-way bounce-back at walls).
//
// TMA cannot wrap, so halo cells that lie across a periodic y/z boundary (only on tiles touching the
// lattice edge) are fetched with ordinary loads; out-of-bounds box elements are zero-filled and unused.
// Physics per cell: sc_cell.cuh.
#include <cuda.h>

#include <cstdlib>

#include "sc_tma.cuh"

namespace clbm {

// CY > 1: the kernel is launched in thread-block clusters of CY tiles along y that cross a cluster barrier once per
// plane.  Nothing is exchanged through it -- it only keeps y-neighbouring tiles on the same plane, so that the halo
// rows they share (box rows y0-1 / y0+TY of one tile are own rows of the next) are fetched from HBM once and hit in
// L2 the second time.  Without it tiles drift apart by several planes and the rows are evicted in between (ncu at
// 512^3: 31.6 GB read for 20.4 GB of populations).
//
// NS stages; plane r of a chunk lives in stage r % NS.  EARLY = 0: a stage is refilled (plane r + NS) once plane r + 1 has arrived
// and the own populations of plane r have been read -- with NS = 2 its box is then in flight only while plane r is collided and
// stored, and a fifth of the warp time is spent waiting for it (profiles/r1_sc_d3q19_512_ncu_full_f.txt: the mbarrier wait holds
// 20 % of the samples).  EARLY >= 1: the own populations of plane r are read FIRST and the stage is handed back before the wait for
// plane r + 1, so that two boxes are in flight while psi of plane r + 1 is built (1: __syncthreads; 2: an "empty" mbarrier the
// warps arrive on, only the issuing thread waits; 3: a shared-memory counter, the LAST warp to arrive issues the refill and nobody
// waits).
//
// SPLIT = 2: the CTA is two independent groups of TY / 2 rows that share nothing but the staged boxes: each has its own psi ring
// (the row next to the other group is recomputed rather than exchanged) and its own named barrier, so that the groups drift
// apart by up to a plane and the shared-memory, FP64 and store phases of one overlap those of the other -- what a second CTA per
// SM would give, without a second set of boxes (two 4 x 64 CTAs would need 2 x 124 KB and re-fetch the rows between them).
//
// IDX32: the whole "out" buffer is addressed with unsigned 32-bit element indices from one base pointer (19 * ncs < 2^32).  The
// general form spends seven integer instructions per store on k * ncs + i + offset in 64 bits (133 of 747 per node in
// profiles/r2_sc_d3q19_512_ncu_full_g.txt); this one two or three.
template <int TY, int TZ, int MINB, int CY, int NS, int EARLY, int SPY, int SPZ, int IDX32>
__global__ void __launch_bounds__(TY *TZ, MINB)
sc_fused_tma_kernel(const __grid_constant__ CUtensorMap tmap, const OutTable P, const uint8_t *__restrict__ flag,
                    const double *__restrict__ fin, const double *__restrict__ psi_g, Geom g, ModelParams mp, int xchunk,
                    int x_begin, int x_end, int nch1, int x2_begin, int x2_end)
{
    using C = TmaCfg<TY, TZ, NS, SPY, SPZ>;
    constexpr int SPLIT = SPY * SPZ;
    static_assert(SPLIT == 1 || EARLY == 3, "independent groups need the waiting-free refill");
    static_assert(SPLIT == 1 || CY == 1, "no lock-step clusters of split CTAs");
    constexpr int NWARP = (TY * TZ + 31) / 32, GY = C::GY, GZ = C::GZ, GT = C::GT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t stage_a = smem_u32(smem_raw);   // shared-window address of stage 0 (the others follow)
    const int tid = threadIdx.x;
    const int gq = SPLIT == 1 ? 0 : tid / GT, lt = SPLIT == 1 ? tid : tid % GT;   // group, thread within the group
    double (*ring)[C::RY][C::RZ] = reinterpret_cast<double (*)[C::RY][C::RZ]>(smem_raw + NS * C::STAGE_BYTES) + gq * 4;
    const int gy0 = (gq / SPZ) * GY, gz0 = (gq % SPZ) * GZ;   // first row / column of the group within the tile
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + NS * C::STAGE_BYTES + C::RING_BYTES);
    uint64_t *mbar_empty = mbar + NS;
    int *refill_cnt = reinterpret_cast<int *>(mbar_empty + NS);

    const int tzl = lt % GZ, tyl = lt / GZ, ty = gy0 + tyl, tz = gz0 + tzl;   // position within the group, within the tile
    const int y0 = blockIdx.y * TY, z0 = blockIdx.x * TZ;
    const int y = y0 + ty, z = z0 + tz;
    const bool inside = (y < g.ny) && (z < g.nz);
    // this CTA collides planes [xa, xa + nplanes) of [x_begin, x_end) -- or, for blockIdx.z >= nch1, of the second range
    // [x2_begin, x2_end) (the two boundary planes of the overlap protocol share one launch)
    const bool second = (int)blockIdx.z >= nch1;
    const int xa = second ? x2_begin + ((int)blockIdx.z - nch1) * xchunk : x_begin + (int)blockIdx.z * xchunk;
    const int nplanes = min(second ? x2_end : x_end, xa + xchunk) - xa;
    const int plane = (int)g.plane, nz = g.nz, G = g.G;
    const int ty_n = max(0, min(GY, g.ny - (y0 + gy0))), tz_n = max(0, min(GZ, g.nz - (z0 + gz0)));   // rows / columns of the group inside the lattice
    const int nrow = tz_n + 2, nhalo = (ty_n > 0 && tz_n > 0) ? 2 * nrow + 2 * ty_n : 0;
    const int yz = y * nz + z;
    const int own_s = (ty + 1) * C::BZ + (tz + 2);   // own cell inside a staged k-slab (box column = ring column + 1)

    // this thread's halo cell (if any): position in the group's ring, in the box, wrapped lattice position, whether TMA
    // could not fetch it
    const bool h_act = lt < nhalo;
    int hsy = 0, hsz = 0;
    if (h_act) {
        if (lt < nrow) { hsy = 0; hsz = lt; }
        else if (lt < 2 * nrow) { hsy = ty_n + 1; hsz = lt - nrow; }
        else { const int q = lt - 2 * nrow; hsy = 1 + (q >> 1); hsz = (q & 1) ? tz_n + 1 : 0; }
    }
    const int hy_raw = y0 + gy0 + hsy - 1, hz_raw = z0 + gz0 + hsz - 1;
    const bool h_wrapped = (hy_raw < 0) || (hy_raw >= g.ny) || (hz_raw < 0) || (hz_raw >= g.nz);
    const int hyz = g.wy(hy_raw) * nz + g.wz(hz_raw);
    const int h_s = (gy0 + hsy) * C::BZ + (gz0 + hsz + 1);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(&mbar[s], 1);
            mbar_init(&mbar_empty[s], NWARP);
            refill_cnt[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // r = 0 .. nplanes+1 enumerates the planes xa-1 .. xa+nplanes; plane r lives in stage r % NS, ring slot r&3
    auto xs_of = [&](int r) { return g.wx(xa - 1 + r) + G; };   // storage plane
    auto issue = [&](int r) {
        mbar_expect_tx(&mbar[r % NS], (uint32_t)(C::BOX * 8));
        tma_load_4d(stage_a + (r % NS) * C::STAGE_BYTES, &tmap, &mbar[r % NS], z0 - 2, y0 - 1, xs_of(r), 0);
    };
    auto wait_full = [&](int r) { mbar_wait(&mbar[r % NS], (uint32_t)((r / NS) & 1)); };
    // this thread is done with the stage of plane r; once every thread is, the stage is refilled with plane r + NS
    auto release_and_refill = [&](int r) {
        if (EARLY == 3) {
            __syncwarp();
            if ((tid & 31) == 0) {
                __threadfence_block();
                const int old = atomicAdd(&refill_cnt[r % NS], 1);
                if (old == NWARP - 1) {
                    // last warp out: the counter is next touched after the refill has landed, so a plain reset is safe
                    refill_cnt[r % NS] = 0;
                    __threadfence_block();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    if (r + NS <= nplanes + 1) issue(r + NS);
                }
            }
        } else if (EARLY == 2) {
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&mbar_empty[r % NS]);
            if (tid == 0 && r + NS <= nplanes + 1) {
                mbar_wait(&mbar_empty[r % NS], (uint32_t)((r / NS) & 1));
                issue(r + NS);
            }
        } else {
            __syncthreads();
            if (tid == 0 && r + NS <= nplanes + 1) issue(r + NS);
        }
    };
    double psn = 0.0, rhn = 0.0;
    bool gpn = true;
    // psi of plane r (rows of the group + halo ring) from its staged box into the group's ring; keeps the own psi / G1 branch
    auto make_psi = [&](int r, uint8_t fl_own, uint8_t fl_halo) {
        const uint32_t st = stage_a + (r % NS) * C::STAGE_BYTES;
        const int xg = xa - 1 + r;
        if (!g.wrapx && (xg < 0 || xg >= g.nx)) {
            // x-slab mode: planes -1 and nx belong to the neighbour slab; their psi arrived with the moment halo
            // exchange (|value| = psi, see sc_psi_kernel) and their mask sits in the ghost planes of flag[]
            const int xs = xg + G;
            if (inside) ring[r & 3][tyl + 1][tzl + 1] = (fl_own == CELL_BB) ? -1.0 : fabs(psi_g[xs * plane + yz]);
            if (h_act) ring[r & 3][hsy][hsz] = (fl_halo == CELL_BB) ? -1.0 : fabs(psi_g[xs * plane + hyz]);
            return;
        }
        if (inside) {
            double v = -1.0;
            psn = 0.0;
            rhn = 0.0;
            gpn = true;
            if (fl_own != CELL_BB) {
                double f[19];
#pragma unroll
                for (int k = 0; k < 19; ++k) f[k] = lds_f64(st + (k * (C::SY * C::BZ) + own_s) * 8);
                rhn = Mom<L3>::sum(f);
                psn = sc_psi_g1(mp, rhn, gpn);
                v = psn;
            }
            ring[r & 3][tyl + 1][tzl + 1] = v;
        }
        if (h_act) {
            double v = -1.0;
            if (fl_halo != CELL_BB) {
                double f[19];
                if (h_wrapped) {
                    const int i = xs_of(r) * plane + hyz;
#pragma unroll
                    for (int k = 0; k < 19; ++k) f[k] = fin[(size_t)k * g.ncs + i];
                } else {
#pragma unroll
                    for (int k = 0; k < 19; ++k) f[k] = lds_f64(st + (k * (C::SY * C::BZ) + h_s) * 8);
                }
                bool gph;
                v = sc_psi_g1(mp, Mom<L3>::sum(f), gph);
            }
            ring[r & 3][hsy][hsz] = v;
        }
    };
    auto flags_of = [&](int r, uint8_t &fo, uint8_t &fh) {
        const int xs = xs_of(r);
        fo = inside ? flag[xs * plane + yz] : CELL_BB;
        fh = h_act ? flag[xs * plane + hyz] : CELL_BB;
    };

    if (tid == 0) {
#pragma unroll
        for (int r = 0; r < NS; ++r)
            if (r <= nplanes + 1) issue(r);
    }
    uint8_t fo, fh;
    // wmask bit (r & 3): plane r has a bounce_back node inside this group's rows + halo ring (group-uniform).  Planes
    // without walls around take the branch-free force sums and unconditional pushes below.
    unsigned wmask = 0;
    auto has_wall = [&](uint8_t fl_own, uint8_t fl_halo) { return (int)((inside && fl_own == CELL_BB) || (h_act && fl_halo == CELL_BB)); };
    flags_of(0, fo, fh);
    int w0 = has_wall(fo, fh);
    wait_full(0);
    make_psi(0, fo, fh);
    if (EARLY) release_and_refill(0);   // plane 0 only feeds psi: its stage goes back at once
    flags_of(1, fo, fh);
    const int w1 = has_wall(fo, fh);
    wait_full(1);
    make_psi(1, fo, fh);
    double psc = psn, rhc = rhn;
    bool gpc = gpn;
    flags_of(2, fo, fh);        // the node mask runs one plane ahead of its use so that its latency never shows
    w0 = group_sync_or<SPLIT, GT>(w0 | (w1 << 1), 1 + gq);
    wmask = (unsigned)w0 & 3u;   // an OR over both prologue planes: conservative (bit set where either plane has a wall)
    if (wmask) wmask = 3u;
    if (!EARLY && tid == 0 && NS <= nplanes + 1) issue(NS);

    const int oym = (g.wy(y - 1) - y) * nz, oyp = (g.wy(y + 1) - y) * nz;
    const int ozm = g.wz(z - 1) - z, ozp = g.wz(z + 1) - z;

    for (int r = 1; r <= nplanes; ++r) {
        const uint8_t fo_now = fo, fh_now = fh;
        if (r + 2 <= nplanes + 1) flags_of(r + 2, fo, fh);
        // own populations of plane r out of its stage before the stage is recycled
        double fc[19];
        auto load_own = [&]() {
            const uint32_t st = stage_a + (r % NS) * C::STAGE_BYTES;
#pragma unroll
            for (int k = 0; k < 19; ++k) fc[k] = lds_f64(st + (k * (C::SY * C::BZ) + own_s) * 8);
        };
        if (EARLY) {
            load_own();
            release_and_refill(r);
        }
        wait_full(r + 1);
        make_psi(r + 1, fo_now, fh_now);
        if (!EARLY) load_own();
        {
            const unsigned bit = 1u << ((r + 1) & 3);
            wmask = group_sync_or<SPLIT, GT>(has_wall(fo_now, fh_now), 1 + gq) ? (wmask | bit) : (wmask & ~bit);
        }
        if (CY > 1) {
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
        if (!EARLY && tid == 0 && r + NS <= nplanes + 1) issue(r + NS);

        const int sm = (r + 3) & 3, s0 = r & 3, sp = (r + 1) & 3;
        const bool walls = (wmask & ((1u << sm) | (1u << s0) | (1u << sp))) != 0u;
        if (!walls) {
            if (inside) {
                // no bounce_back node within reach: every neighbour is fluid, every push goes to the neighbour
                ScForceSums s = {{0., 0., 0.}, {0., 0., 0.}, 0u};
#pragma unroll
                for (int k = 0; k < 19; ++k) {
                    if (k == L3::REST) continue;
                    const int slot = L3::cx(k) < 0 ? sm : (L3::cx(k) > 0 ? sp : s0);
                    const double v = ring[slot][tyl + 1 + L3::cy(k)][tzl + 1 + L3::cz(k)];
                    if (L3::cx(k)) s.ff[0] += L3::t(k) * L3::cx(k) * v;
                    if (L3::cy(k)) s.ff[1] += L3::t(k) * L3::cy(k) * v;
                    if (L3::cz(k)) s.ff[2] += L3::t(k) * L3::cz(k) * v;
                }
                double out[19];
                sc_collide_rho<L3>(mp, fc, s, rhc, psc, gpc, out);
                const int x = xa - 1 + r;
                const int i = (x + G) * plane + yz;
                const int oxm = (g.wx(x - 1) - x) * plane, oxp = (g.wx(x + 1) - x) * plane;
                if (IDX32) {
                    const unsigned i0 = (unsigned)i, im = (unsigned)(i + oxm), ip = (unsigned)(i + oxp);
#pragma unroll
                    for (int k = 0; k < 19; ++k) {
                        unsigned idx = (L3::cx(k) < 0 ? im : (L3::cx(k) > 0 ? ip : i0)) + P.kn[k];
                        if (L3::cy(k)) idx += (unsigned)(L3::cy(k) < 0 ? oym : oyp);
                        if (L3::cz(k)) idx += (unsigned)(L3::cz(k) < 0 ? ozm : ozp);
                        P.base[idx] = out[k];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 19; ++k) {
                        const int off = (L3::cx(k) < 0 ? oxm : (L3::cx(k) > 0 ? oxp : 0)) + (L3::cy(k) < 0 ? oym : (L3::cy(k) > 0 ? oyp : 0)) +
                                        (L3::cz(k) < 0 ? ozm : (L3::cz(k) > 0 ? ozp : 0));
                        P.at(k)[i + off] = out[k];
                    }
                }
            }
        } else if (inside && ring[s0][tyl + 1][tzl + 1] >= 0.0) {
            ScForceSums s = {{0., 0., 0.}, {0., 0., 0.}, 0u};
#pragma unroll
            for (int k = 0; k < 19; ++k) {
                if (k == L3::REST) continue;
                const int slot = L3::cx(k) < 0 ? sm : (L3::cx(k) > 0 ? sp : s0);
                const double v = ring[slot][tyl + 1 + L3::cy(k)][tzl + 1 + L3::cz(k)];
                sc_force_add<L3>(s, k, v < 0.0, v);
            }
            double out[19];
            sc_collide_rho<L3>(mp, fc, s, rhc, psc, gpc, out);

            const int x = xa - 1 + r;
            const int xp = g.wx(x + 1), xm = g.wx(x - 1);
            const int i = (x + G) * plane + yz;
            const int oxm = (xm - x) * plane, oxp = (xp - x) * plane;
#pragma unroll
            for (int k = 0; k < 19; ++k) {
                if (k == L3::REST) { P.at(k)[i] = out[k]; continue; }
                const int off = (L3::cx(k) < 0 ? oxm : (L3::cx(k) > 0 ? oxp : 0)) + (L3::cy(k) < 0 ? oym : (L3::cy(k) > 0 ? oyp : 0)) +
                                (L3::cz(k) < 0 ? ozm : (L3::cz(k) > 0 ? ozp : 0));
                if (s.wall & (1u << k)) P.at(L3::opp(k))[i] = out[k];
                else P.at(k)[i + off] = out[k];
            }
        }
        psc = psn;
        rhc = rhn;
        gpc = gpn;
    }
}

// can this lattice be described by a 16-byte-stride tensor map?
bool sc_tma_eligible(const clbm_ctx *c)
{
    const Geom &g = c->geo;
    return c->Q == 19 && (g.nz % 2 == 0) && g.ncs < (1LL << 31) && get_encode() != nullptr;
}

template <int TY, int TZ, int MINB, int CY, int NS = 2, int EARLY = 0, int SPY = 1, int SPZ = 1, int IDX32 = 0>
static int launch_tma_c(clbm_ctx *c, int x_begin, int x_end, int x2_begin, int x2_end)
{
    if (IDX32 && 19ull * (unsigned long long)c->geo.ncs >= (1ull << 32))   // too large for 32-bit element indices: the general form
        return launch_tma_c<TY, TZ, MINB, CY, NS, EARLY, SPY, SPZ, 0>(c, x_begin, x_end, x2_begin, x2_end);
    using C = TmaCfg<TY, TZ, NS, SPY, SPZ>;
    static_assert(C::SMEM <= 232448, "stages + psi ring must fit the 227 KB a CTA may opt in to");
    const Geom &g = c->geo;
    CUtensorMap tmap;
    const cuuint32_t box[4] = {(cuuint32_t)C::BZ, (cuuint32_t)C::SY, 1, 19};
    // box rows are 544 B starting 16 B before a 512-B boundary: 256-B promotion over-fetches a third of every row
    // (ncu at 512^3: 31.6 GB read for 20.4 GB of populations); 64 B measured best (profiles/README.md)
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
    if (c->env.tma_promo >= 0) {
        const int v = c->env.tma_promo;
        promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : (v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : promo));
    }
    if (int rc = cached_tmap(c, c->pop[0][c->parity], box, promo, &tmap)) return rc;

    const int tiles = ((g.ny + TY - 1) / TY) * ((g.nz + TZ - 1) / TZ);
    const int nxr = x_end - x_begin;
    if (nxr <= 0) return 0;
    // Short x-chunks keep the CTAs that are resident at the same time on neighbouring planes, so the halo rows that
    // y/z-neighbouring tiles share are still in L2 when the second tile asks for them: at 512^3 chunks of 20-24 planes
    // give 16.3 GLUPS against 14.8 for 256-plane chunks, although every chunk re-reads two prologue planes (sweep in
    // profiles/README.md).  (tiles is only used to keep at least one full wave of CTAs.)
    int xchunk = nxr < 24 ? nxr : 24;
    if ((long long)tiles * ((nxr + xchunk - 1) / xchunk) < 148LL * MINB && nxr > 8) xchunk = 8;
    if (c->env.sc_xchunk > 0) xchunk = c->env.sc_xchunk < nxr ? c->env.sc_xchunk : nxr;
    const int nch1 = (nxr + xchunk - 1) / xchunk, nch2 = x2_end > x2_begin ? (x2_end - x2_begin + xchunk - 1) / xchunk : 0;
    dim3 grid((g.nz + TZ - 1) / TZ, (g.ny + TY - 1) / TY, nch1 + nch2);
    OutTable P = {c->pop[0][1 - c->parity], (size_t)g.ncs, {0}};
    for (int k = 0; k < 19; ++k) P.kn[k] = (unsigned)((unsigned long long)k * (unsigned long long)g.ncs);
    auto kern = sc_fused_tma_kernel<TY, TZ, MINB, CY, NS, EARLY, SPY, SPZ, IDX32>;
    static PerDeviceOnce attr;
    if (attr.need(c->device)) {
        CLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr.mark(c->device);
    }
    LaunchScope ls(c, "sc_fused_tma_collide_stream", nxr * 2 >= g.nx);   // the boundary-plane launches of the overlap protocol are not the dominant kernel
    if (CY == 1) {
        kern<<<grid, TY * TZ, C::SMEM, c->stream>>>(tmap, P, c->flag, c->pop[0][c->parity], c->fld[0], g, c->mp, xchunk, x_begin, x_end, nch1, x2_begin, x2_end);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(TY * TZ, 1, 1);
        cfg.dynamicSmemBytes = C::SMEM;
        cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 1;
        at[0].val.clusterDim.y = CY;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const uint8_t *fl = c->flag;
        const double *fin = c->pop[0][c->parity], *psi = c->fld[0];
        CLBM_CUDA(cudaLaunchKernelEx(&cfg, kern, tmap, P, fl, fin, psi, g, c->mp, xchunk, x_begin, x_end, nch1, x2_begin, x2_end));
    }
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

// lock-step clusters along y are an experiment (CLBM_SC_CLUSTER = 2 / 4): measured SLOWER at 512^3 (12.8 / 11.8 vs 14.8
// GLUPS) -- waiting for the slower partner costs more than the shared halo rows save -- so the default is 1
template <int TY, int TZ, int MINB, int NS = 2, int EARLY = 0, int SPY = 1, int SPZ = 1, int IDX32 = 0>
static int launch_tma(clbm_ctx *c, int x_begin, int x_end, int x2_begin, int x2_end)
{
    if (NS != 2 || EARLY != 0) return launch_tma_c<TY, TZ, MINB, 1, NS, EARLY, SPY, SPZ, IDX32>(c, x_begin, x_end, x2_begin, x2_end);
    const int cy = c->env.sc_cluster > 0 ? c->env.sc_cluster : 1;
    const int ytiles = (c->geo.ny + TY - 1) / TY;
    if (MINB == 1 && cy >= 4 && ytiles % 4 == 0) return launch_tma_c<TY, TZ, MINB, 4>(c, x_begin, x_end, x2_begin, x2_end);
    if (MINB == 1 && cy >= 2 && ytiles % 2 == 0) return launch_tma_c<TY, TZ, MINB, 2>(c, x_begin, x_end, x2_begin, x2_end);
    return launch_tma_c<TY, TZ, MINB, 1>(c, x_begin, x_end, x2_begin, x2_end);
}

// collide + push of the planes [x_begin, x_end) and [x2_begin, x2_end) of the slab in one launch
int sc_fused_tma_range(clbm_ctx *c, int variant, int x_begin, int x_end, int x2_begin, int x2_end)
{
    int rc;
    switch (variant) {
    case 11: rc = launch_tma<8, 64, 1>(c, x_begin, x_end, x2_begin, x2_end); break;
    case 12: rc = launch_tma<16, 32, 1>(c, x_begin, x_end, x2_begin, x2_end); break;
    case 13: rc = launch_tma<4, 64, 1>(c, x_begin, x_end, x2_begin, x2_end); break;
    case 14: rc = launch_tma<4, 32, 3>(c, x_begin, x_end, x2_begin, x2_end); break;
    case 15: rc = launch_tma<8, 16, 3>(c, x_begin, x_end, x2_begin, x2_end); break;
    case 16: rc = launch_tma<8, 32, 1>(c, x_begin, x_end, x2_begin, x2_end); break;
    case 21: rc = launch_tma<8, 64, 1, 2, 1>(c, x_begin, x_end, x2_begin, x2_end); break;
    case 22: rc = launch_tma<8, 64, 1, 2, 2>(c, x_begin, x_end, x2_begin, x2_end); break;
    case 23: rc = launch_tma<8, 64, 1, 2, 3>(c, x_begin, x_end, x2_begin, x2_end); break;
    case 24: rc = launch_tma<8, 64, 1, 2, 3, 2>(c, x_begin, x_end, x2_begin, x2_end); break;
    case 25: rc = launch_tma<8, 64, 1, 2, 3, 1, 2>(c, x_begin, x_end, x2_begin, x2_end); break;
    case 26: rc = launch_tma<8, 64, 1, 2, 3, 1, 4>(c, x_begin, x_end, x2_begin, x2_end); break;
    case 29: rc = launch_tma<8, 64, 1, 2, 3, 2, 1, 1>(c, x_begin, x_end, x2_begin, x2_end); break;
    case 27: rc = launch_tma<16, 32, 1, 2, 3, 2, 2>(c, x_begin, x_end, x2_begin, x2_end); break;
    case 28: rc = launch_tma<16, 32, 1, 2, 3, 4, 1>(c, x_begin, x_end, x2_begin, x2_end); break;
    default: rc = launch_tma<6, 32, 2>(c, x_begin, x_end, x2_begin, x2_end); break;
    }
    return rc;
}

int sc_fused_tma_step(clbm_ctx *c, int variant) { return sc_fused_tma_range(c, variant, 0, c->geo.nx, 0, 0); }

}  // namespace clbm
