// sc_fused_tma_persist.cu -- the TMA-staged D3Q19 Shan-Chen step (sc_fused_tma.cu) as a PERSISTENT kernel.
//
// sc_fused_tma_kernel launches one CTA per (tile, x-chunk): 11 264 CTAs at 512^3, one resident per SM.  Every CTA starts with
// an empty pipeline -- mbarrier set-up, two TMA boxes requested, nothing to do until the first lands, two psi-only planes -- and
// nothing else runs on its SM meanwhile: 7 % of the warp samples of profiles/r2_sc_d3q19_512_ncu_full_g.txt sit in that
// prologue.  Here one CTA per SM stays resident and takes (tile, x-chunk) work items from a queue (an atomic counter, same
// order as the grid of the other kernel, so that CTAs resident at the same time still work on neighbouring tiles and share
// their halo rows through L2).  The box pipeline never drains: the stage released by plane n of one item is refilled with
// plane 0 of the next item, whose id the CTA fetched when it started the current one.  What is left of the per-item cost is
// the re-computation of the thread geometry and two psi-only planes.
//
// Everything else -- stages, early release by the last warp, independent row groups with their own psi rings, 32-bit store
// indices, wall-free fast path -- is the same as variant 29 of sc_fused_tma.cu; the arithmetic per node is identical
// (bit-identical populations, tools/sc3d_variants.py).
#include <cuda.h>

#include <cstdlib>

#include "sc_tma.cuh"

namespace clbm {

int sc_persist_queue(clbm_ctx *c, int **q);   // sc_fused.cu

// what a work item is made of (uniform; cheap enough to recompute from the item number wherever it is needed)
struct ItemGeo {
    int xa, nplanes, y0, z0;
};

struct PersistArgs {
    int xchunk, x_begin, x_end, nch1, x2_begin, x2_end;
    int ytiles, ztiles, nitems;
    int static_sched;   // 1: item = blockIdx.x + it * gridDim.x instead of the queue (CLBM_SC_PERSIST_STATIC, an experiment)
    int *queue;   // [0]: next item, [1]: CTAs that have finished (the last one resets both)
};

template <int TY, int TZ>
CLBM_D ItemGeo item_geo(const PersistArgs &A, int item)
{
    const int tiles = A.ytiles * A.ztiles;
    const int bz = item / tiles, t = item - bz * tiles;
    const int by = t / A.ztiles, bx = t - by * A.ztiles;
    ItemGeo q;
    const bool second = bz >= A.nch1;
    q.xa = second ? A.x2_begin + (bz - A.nch1) * A.xchunk : A.x_begin + bz * A.xchunk;
    q.nplanes = min(second ? A.x2_end : A.x_end, q.xa + A.xchunk) - q.xa;
    q.y0 = by * TY;
    q.z0 = bx * TZ;
    return q;
}

template <int TY, int TZ, int NS, int SPY, int SPZ, int IDX32>
__global__ void __launch_bounds__(TY *TZ, 1)
sc_fused_tma_persist_kernel(const __grid_constant__ CUtensorMap tmap, const OutTable P, const uint8_t *__restrict__ flag,
                            const double *__restrict__ fin, const double *__restrict__ psi_g, Geom g, ModelParams mp, PersistArgs A)
{
    using C = TmaCfg<TY, TZ, NS, SPY, SPZ>;
    constexpr int SPLIT = SPY * SPZ;
    constexpr int NWARP = (TY * TZ + 31) / 32, GY = C::GY, GZ = C::GZ, GT = C::GT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t stage_a = smem_u32(smem_raw);
    const int tid = threadIdx.x;
    const int gq = SPLIT == 1 ? 0 : tid / GT, lt = SPLIT == 1 ? tid : tid % GT;
    double (*ring)[C::RY][C::RZ] = reinterpret_cast<double (*)[C::RY][C::RZ]>(smem_raw + NS * C::STAGE_BYTES) + gq * 4;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + NS * C::STAGE_BYTES + C::RING_BYTES);
    int *refill_cnt = reinterpret_cast<int *>(mbar + NS);
    volatile int *s_item = reinterpret_cast<volatile int *>(refill_cnt + NS);   // [0], [1]: item of iteration it & 1 ... see below
    const int gy0 = (gq / SPZ) * GY, gz0 = (gq % SPZ) * GZ;
    const int tzl = lt % GZ, tyl = lt / GZ, ty = gy0 + tyl, tz = gz0 + tzl;
    const int plane = (int)g.plane, nz = g.nz, G = g.G;
    const int own_s = (ty + 1) * C::BZ + (tz + 2);

    // box of plane r of work item `item` into the stage of sequence number seq
    auto issue = [&](int item, int r, int seq) {
        const ItemGeo q = item_geo<TY, TZ>(A, item);
        mbar_expect_tx(&mbar[seq % NS], (uint32_t)(C::BOX * 8));
        tma_load_4d(stage_a + (seq % NS) * C::STAGE_BYTES, &tmap, &mbar[seq % NS], q.z0 - 2, q.y0 - 1, g.wx(q.xa - 1 + r) + G, 0);
    };
    auto wait_full = [&](int seq) { mbar_wait(&mbar[seq % NS], (uint32_t)((seq / NS) & 1)); };

    // s_item[it & 1] = the work item of the CTA's it-th iteration.  Item it + 1 is fetched by the last warp to release plane
    // nplanes - 2 of item it; the value is read (a) by whichever warp is the last to release one of the final NS stages of item it --
    // after an atomic on a release counter that the fetching warp also took part in, behind a fence -- and (b) by every thread when
    // it moves on to item it + 1 (behind the mbarrier wait for a box that was requested after (a)).
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(&mbar[s], 1);
            refill_cnt[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int first = A.static_sched ? (int)blockIdx.x : atomicAdd(&A.queue[0], 1);
        s_item[0] = first;
        if (first < A.nitems) {
            const ItemGeo q = item_geo<TY, TZ>(A, first);
#pragma unroll
            for (int r = 0; r < NS; ++r)
                if (r <= q.nplanes + 1) issue(first, r, r);   // nplanes >= 1 and NS = 2: always
        }
    }
    __syncthreads();

    int seq0 = 0;   // sequence number of plane 0 of the current item (the same in every thread)
    for (int it = 0;; ++it) {
        const int item = s_item[it & 1];
        if (item >= A.nitems) break;
        const ItemGeo Q = item_geo<TY, TZ>(A, item);
        const int xa = Q.xa, nplanes = Q.nplanes, y0 = Q.y0, z0 = Q.z0;
        const int y = y0 + ty, z = z0 + tz;
        const bool inside = (y < g.ny) && (z < g.nz);
        const int ty_n = max(0, min(GY, g.ny - (y0 + gy0))), tz_n = max(0, min(GZ, g.nz - (z0 + gz0)));
        const int nrow = tz_n + 2, nhalo = (ty_n > 0 && tz_n > 0) ? 2 * nrow + 2 * ty_n : 0;
        const int yz = y * nz + z;
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
        const int oym = (g.wy(y - 1) - y) * nz, oyp = (g.wy(y + 1) - y) * nz;
        const int ozm = g.wz(z - 1) - z, ozp = g.wz(z + 1) - z;

        auto xs_of = [&](int r) { return g.wx(xa - 1 + r) + G; };
        // this thread is done with the stage of plane r of the item; the last warp to say so refills it with the plane NS further
        // on -- of this item, or of the next one
        auto release_and_refill = [&](int r) {
            __syncwarp();
            if ((tid & 31) == 0) {
                const int seq = seq0 + r;
                __threadfence_block();
                const int old = atomicAdd(&refill_cnt[seq % NS], 1);
                if (old == NWARP - 1) {
                    refill_cnt[seq % NS] = 0;   // next touched after the refill has landed
                    __threadfence_block();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    // the NEXT item is taken from the queue as late as possible (two planes before its first box is requested): CTAs
                    // that hold a reservation for a whole item work on tiles whose neighbours were done an item ago, and the halo
                    // rows are no longer in L2 (31.3 GB read instead of 23.7 GB when the item was fetched at the start of the current one)
                    if (r == max(0, nplanes - 2)) {
                        s_item[(it + 1) & 1] = A.static_sched ? item + (int)gridDim.x : atomicAdd(&A.queue[0], 1);
                        __threadfence_block();
                    }
                    if (r + NS <= nplanes + 1) issue(item, r + NS, seq + NS);
                    else {
                        const int nxt = s_item[(it + 1) & 1];
                        if (nxt < A.nitems) issue(nxt, r + NS - (nplanes + 2), seq + NS);
                    }
                }
            }
        };
        double psn = 0.0, rhn = 0.0;
        bool gpn = true;
        auto make_psi = [&](int r, uint8_t fl_own, uint8_t fl_halo) {
            const int seq = seq0 + r;
            const uint32_t st = stage_a + (seq % NS) * C::STAGE_BYTES;
            const int xg = xa - 1 + r;
            if (!g.wrapx && (xg < 0 || xg >= g.nx)) {
                const int xs = xg + G;
                if (inside) ring[seq & 3][tyl + 1][tzl + 1] = (fl_own == CELL_BB) ? -1.0 : fabs(psi_g[xs * plane + yz]);
                if (h_act) ring[seq & 3][hsy][hsz] = (fl_halo == CELL_BB) ? -1.0 : fabs(psi_g[xs * plane + hyz]);
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
                ring[seq & 3][tyl + 1][tzl + 1] = v;
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
                ring[seq & 3][hsy][hsz] = v;
            }
        };
        auto flags_of = [&](int r, uint8_t &fo, uint8_t &fh) {
            const int xs = xs_of(r);
            fo = inside ? flag[xs * plane + yz] : CELL_BB;
            fh = h_act ? flag[xs * plane + hyz] : CELL_BB;
        };
        auto has_wall = [&](uint8_t fl_own, uint8_t fl_halo) { return (int)((inside && fl_own == CELL_BB) || (h_act && fl_halo == CELL_BB)); };

        // the masks of the first three planes at once (one memory latency per item; later planes run two ahead of their use)
        uint8_t fo0, fh0, fo1, fh1, fo, fh;
        flags_of(0, fo0, fh0);
        flags_of(1, fo1, fh1);
        flags_of(2, fo, fh);
        unsigned wmask = 0;   // bit (seq & 3): that plane has a bounce_back node inside this group's rows + halo ring
        wait_full(seq0);
        make_psi(0, fo0, fh0);
        release_and_refill(0);   // plane 0 only feeds psi: its stage goes back at once
        // a barrier between the last collision of the previous item (which read the ring slot psi of plane 1 goes into) and psi(1)
        if (group_sync_or<SPLIT, GT>(has_wall(fo0, fh0), 1 + gq)) wmask |= 1u << (seq0 & 3);
        wait_full(seq0 + 1);
        make_psi(1, fo1, fh1);
        double psc = psn, rhc = rhn;
        bool gpc = gpn;
        if (group_sync_or<SPLIT, GT>(has_wall(fo1, fh1), 1 + gq)) wmask |= 1u << ((seq0 + 1) & 3);

        for (int r = 1; r <= nplanes; ++r) {
            const uint8_t fo_now = fo, fh_now = fh;
            if (r + 2 <= nplanes + 1) flags_of(r + 2, fo, fh);
            double fc[19];
            {
                const uint32_t st = stage_a + ((seq0 + r) % NS) * C::STAGE_BYTES;
#pragma unroll
                for (int k = 0; k < 19; ++k) fc[k] = lds_f64(st + (k * (C::SY * C::BZ) + own_s) * 8);
            }
            release_and_refill(r);
            wait_full(seq0 + r + 1);
            make_psi(r + 1, fo_now, fh_now);
            {
                const unsigned bit = 1u << ((seq0 + r + 1) & 3);
                wmask = group_sync_or<SPLIT, GT>(has_wall(fo_now, fh_now), 1 + gq) ? (wmask | bit) : (wmask & ~bit);
            }
            const int sm = (seq0 + r + 3) & 3, s0 = (seq0 + r) & 3, sp = (seq0 + r + 1) & 3;
            const bool walls = (wmask & ((1u << sm) | (1u << s0) | (1u << sp))) != 0u;
            if (!walls) {
                if (inside) {
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
                const int i = (x + G) * plane + yz;
                const int oxm = (g.wx(x - 1) - x) * plane, oxp = (g.wx(x + 1) - x) * plane;
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
        // plane nplanes + 1 only fed psi: hand its stage back (every thread of the group is past its reads: the barrier above)
        release_and_refill(nplanes + 1);
        seq0 += nplanes + 2;
    }

    // the last CTA to leave resets the queue for the next launch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&A.queue[1], 1) == (int)gridDim.x - 1) {
            A.queue[0] = 0;
            A.queue[1] = 0;
            __threadfence();
        }
    }
}

template <int TY, int TZ, int NS, int SPY, int SPZ>
static int launch_persist(clbm_ctx *c, int x_begin, int x_end, int x2_begin, int x2_end)
{
    using C = TmaCfg<TY, TZ, NS, SPY, SPZ>;
    static_assert(C::SMEM <= 232448, "stages + psi rings must fit the 227 KB a CTA may opt in to");
    const Geom &g = c->geo;
    CUtensorMap tmap;
    const cuuint32_t box[4] = {(cuuint32_t)C::BZ, (cuuint32_t)C::SY, 1, 19};
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
    if (c->env.tma_promo >= 0) {
        const int v = c->env.tma_promo;
        promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : (v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : promo));
    }
    if (int rc = cached_tmap(c, c->pop[0][c->parity], box, promo, &tmap)) return rc;
    const int nxr = x_end - x_begin;
    if (nxr <= 0) return 0;
    PersistArgs A;
    A.ytiles = (g.ny + TY - 1) / TY;
    A.ztiles = (g.nz + TZ - 1) / TZ;
    const int tiles = A.ytiles * A.ztiles;
    int xchunk = nxr < 24 ? nxr : 24;
    if ((long long)tiles * ((nxr + xchunk - 1) / xchunk) < 148LL && nxr > 8) xchunk = 8;
    if (c->env.sc_xchunk > 0) xchunk = c->env.sc_xchunk < nxr ? c->env.sc_xchunk : nxr;
    A.xchunk = xchunk;
    A.x_begin = x_begin; A.x_end = x_end; A.x2_begin = x2_begin; A.x2_end = x2_end;
    A.nch1 = (nxr + xchunk - 1) / xchunk;
    const int nch2 = x2_end > x2_begin ? (x2_end - x2_begin + xchunk - 1) / xchunk : 0;
    A.nitems = tiles * (A.nch1 + nch2);
    { const char *e = getenv("CLBM_SC_PERSIST_STATIC"); A.static_sched = (e && atoi(e) > 0) ? 1 : 0; }
    if (int rc = sc_persist_queue(c, &A.queue)) return rc;
    OutTable P = {c->pop[0][1 - c->parity], (size_t)g.ncs, {0}};
    for (int k = 0; k < 19; ++k) P.kn[k] = (unsigned)((unsigned long long)k * (unsigned long long)g.ncs);
    const bool idx32 = 19ull * (unsigned long long)g.ncs < (1ull << 32);
    auto k32 = sc_fused_tma_persist_kernel<TY, TZ, NS, SPY, SPZ, 1>;
    auto k64 = sc_fused_tma_persist_kernel<TY, TZ, NS, SPY, SPZ, 0>;
    static PerDeviceOnce attr;
    if (attr.need(c->device)) {
        CLBM_CUDA(cudaFuncSetAttribute(k32, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        CLBM_CUDA(cudaFuncSetAttribute(k64, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr.mark(c->device);
    }
    int grid = c->sm_count > 0 ? c->sm_count : 148;
    if (grid > A.nitems) grid = A.nitems;
    LaunchScope ls(c, "sc_fused_tma_collide_stream", nxr * 2 >= g.nx);
    if (idx32) k32<<<grid, TY * TZ, C::SMEM, c->stream>>>(tmap, P, c->flag, c->pop[0][c->parity], c->fld[0], g, c->mp, A);
    else k64<<<grid, TY * TZ, C::SMEM, c->stream>>>(tmap, P, c->flag, c->pop[0][c->parity], c->fld[0], g, c->mp, A);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

int sc_fused_tma_persist_range(clbm_ctx *c, int variant, int x_begin, int x_end, int x2_begin, int x2_end)
{
    (void)variant;
    return launch_persist<8, 64, 2, 2, 1>(c, x_begin, x_end, x2_begin, x2_end);
}

}  // namespace clbm
