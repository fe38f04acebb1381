// sc2d_tma.cu -- D2Q9 Shan-Chen fused step with TMA-staged population columns (sm_100a): the recipe of sc_fused_tma.cu for the
// 2-D lattice.
//
// The register-pipelined D2Q9 kernel (sc_fused.cu) consumes its loads at once -- psi needs the density -- so every column pays a
// memory latency that 16 warps per SM cannot hide: DRAM at 47 % at 8192^2 (profiles/r2_sc_d2q9_8192_ncu_full_a.txt), 0.62 of the
// measured peak with an L2 prefetch.  Here the loads leave the instruction stream:
//
//   * a CTA owns T consecutive rows (y is the fastest index of a 2-D lattice) and marches along x; one cp.async.bulk.tensor per
//     column fetches the 9 x (T + 4) box "rows y0-2 .. y0+T+1, all directions" into one of NS shared-memory stages (the box starts
//     two rows early because the innermost TMA coordinate must be 16-byte aligned; the halo is ONE row on either side, so
//     the over-fetch is 4 / T);
//   * psi of column x+1 (rows + the two halo rows) is computed from its box into a 4-slot ring of psi columns; column x is collided
//     from its stage and pushed with coalesced stores;
//   * a stage is handed back as soon as the own populations of its column are in registers; the LAST warp to do so issues the
//     refill (no barrier, nobody waits), so NS - 1 boxes are in flight while a column is worked on;
//   * several CTAs per SM (a stage is 9 (T + 4) doubles: 18.7 KB at T = 256), so the phases of one CTA overlap those of the others.
//
// Per-cell arithmetic and summation order are those of sc_fused_kernel<D2Q9>: the populations are BIT-IDENTICAL to it
// (tests/test_gpu_zr_sc2d_tma.py), which is why x-slabs and the MRT operator can stay on that kernel.
// Single slab, BGK, every force variant (Yuan-CS, constant G, Rayleigh-Taylor / Guo), ny even.
#include <cuda.h>

#include <cstdlib>

#include "sc_cell.cuh"
#include "tma.cuh"

namespace clbm {

using L2 = D2Q9;

struct OutTable2 {
    double *base;
    unsigned kn[9];   // k * ncs as 32-bit element offsets (9 * ncs < 2^32)
};

template <int T, int NS>
struct Tma2Cfg {
    static constexpr int BY = T + 4;                                   // box rows y0-2 .. y0+T+1
    static constexpr int BOX = 9 * BY;
    static constexpr int STAGE_BYTES = ((BOX * 8 + 127) / 128) * 128;
    static constexpr int RING_BYTES = 4 * (T + 2) * 8;
    static constexpr int SMEM = NS * STAGE_BYTES + RING_BYTES + 128;   // NS mbarriers + NS counters behind the ring
    static_assert(T % 32 == 0 && NS >= 2 && NS <= 8, "whole warps, 2 to 8 stages");
    // a TMA box dimension holds at most 256 elements: a column box of up to 256 rows is described as it is -- (BY rows, 1 column, 9
    // directions), innermost extent BY * 8 bytes -- and a taller one with its rows as PAIRS (2, BY / 2, 1, 9).  The pair form makes
    // the innermost extent 16 bytes, and the TMA unit fetches a box row by row of that extent: measured on the 128-row default,
    // 17 % of the stall samples sat on the mbarrier wait for a box issued two columns (6 us) earlier (cached_tmap_2d).
    static constexpr bool PAIRS = BY > 256;
    static_assert(BY / 2 <= 256, "a TMA box dimension holds at most 256 elements");
};

// GUO = true: the Rayleigh-Taylor variant (psi = 1 - exp(-rho), a wall neighbour contributes the psi of the opposite neighbour, Guo
// forcing: SC/apps/RayleighTaylor2D.h), a compile-time flag exactly as in sc_fused_kernel.
template <int T, int NS, int MINB, bool GUO>
__global__ void __launch_bounds__(T, MINB)
sc2d_tma_kernel(const __grid_constant__ CUtensorMap tmap, const OutTable2 P, const uint8_t *__restrict__ flag,
                const double *__restrict__ fin, Geom g, ModelParams mp, int xchunk)
{
    using C = Tma2Cfg<T, NS>;
    constexpr int NWARP = T / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t stage_a = smem_u32(smem_raw);
    double (*ring)[T + 2] = reinterpret_cast<double (*)[T + 2]>(smem_raw + NS * C::STAGE_BYTES);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + NS * C::STAGE_BYTES + C::RING_BYTES);
    int *refill_cnt = reinterpret_cast<int *>(mbar + NS);

    const int tid = threadIdx.x;
    const int ny = g.ny, G = g.G;
    const int y0 = blockIdx.x * T, y = y0 + tid;
    const bool inside = y < ny;
    const int t_n = min(T, ny - y0);                       // rows of this tile inside the lattice
    const int xa = blockIdx.y * xchunk;
    const int nplanes = min(g.nx, xa + xchunk) - xa;
    // halo rows: thread 0 owns row y0-1 (ring index 0), thread 1 row y0+t_n (ring index t_n+1)
    const bool h_act = tid < 2;
    const int hy_raw = tid == 0 ? y0 - 1 : y0 + t_n;
    const bool h_wrapped = hy_raw < 0 || hy_raw >= ny;     // TMA cannot wrap: plain loads at the periodic edge
    const int hy = g.wy(hy_raw);
    const int h_ring = tid == 0 ? 0 : t_n + 1;
    const int h_box = tid == 0 ? 1 : t_n + 2;
    const int own_box = tid + 2;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(&mbar[s], 1);
            refill_cnt[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // r = 0 .. nplanes+1 enumerates the columns xa-1 .. xa+nplanes; column r lives in stage r % NS, ring slot r & 3
    auto xs_of = [&](int r) { return g.wx(xa - 1 + r) + G; };
    auto issue = [&](int r) {
        mbar_expect_tx(&mbar[r % NS], (uint32_t)(C::BOX * 8));
        if constexpr (C::PAIRS) tma_load_4d(stage_a + (r % NS) * C::STAGE_BYTES, &tmap, &mbar[r % NS], 0, (y0 - 2) / 2, xs_of(r), 0);   // y0 - 2 is even, also for tile 0 (-2 / 2 = -1)
        else tma_load_3d(stage_a + (r % NS) * C::STAGE_BYTES, &tmap, &mbar[r % NS], y0 - 2, xs_of(r), 0);
    };
    auto wait_full = [&](int r) { mbar_wait(&mbar[r % NS], (uint32_t)((r / NS) & 1)); };
    auto release_and_refill = [&](int r) {
        __syncwarp();
        if ((tid & 31) == 0) {
            __threadfence_block();
            const int old = atomicAdd(&refill_cnt[r % NS], 1);
            if (old == NWARP - 1) {
                refill_cnt[r % NS] = 0;
                __threadfence_block();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (r + NS <= nplanes + 1) issue(r + NS);
            }
        }
    };
    double psn = 0.0;
    bool gpn = true;
    // node masks of the column whose psi is built next, fetched ONE COLUMN AHEAD: the populations arrive through TMA, and a mask
    // byte loaded at its point of use was the only global load left in the loop -- every column paid its latency (ncu at 8192^2:
    // 27 % of the stall samples on the compare behind that load, 25 % at the barrier waiting for the warps stalled there)
    uint8_t fl_n = CELL_BULK, flh_n = CELL_BULK;
    auto fetch_flags = [&](int r) {
        if (r > nplanes + 1) return;
        const int xs = xs_of(r);
        if (inside) fl_n = flag[xs * ny + y];
        if (h_act) flh_n = flag[xs * ny + hy];
    };
    // psi of column r (own rows + the two halo rows) from its staged box into the ring; keeps the own psi / G1 branch
    auto make_psi = [&](int r) {
        const uint32_t st = stage_a + (r % NS) * C::STAGE_BYTES;
        const int xs = xs_of(r);
        const uint8_t fl_own = fl_n, fl_halo = flh_n;
        fetch_flags(r + 1);
        if (inside) {
            double v = -1.0;
            psn = 0.0;
            gpn = true;
            if (fl_own != CELL_BB) {
                double f[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) f[k] = lds_f64(st + (k * C::BY + own_box) * 8);
                if constexpr (GUO) psn = scrt_psi(Mom<L2>::sum(f));
                else psn = sc_psi_g1(mp, Mom<L2>::sum(f), gpn);
                v = psn;
            }
            ring[r & 3][tid + 1] = v;
        }
        if (h_act) {
            double v = -1.0;
            const int i = xs * ny + hy;
            if (fl_halo != CELL_BB) {
                double f[9];
                if (h_wrapped) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) f[k] = fin[(size_t)k * g.ncs + i];
                } else {
#pragma unroll
                    for (int k = 0; k < 9; ++k) f[k] = lds_f64(st + (k * C::BY + h_box) * 8);
                }
                bool gph;
                if constexpr (GUO) v = scrt_psi(Mom<L2>::sum(f));
                else v = sc_psi_g1(mp, Mom<L2>::sum(f), gph);
            }
            ring[r & 3][h_ring] = v;
        }
    };

    if (tid == 0) {
#pragma unroll
        for (int r = 0; r < NS; ++r)
            if (r <= nplanes + 1) issue(r);
    }
    fetch_flags(0);
    wait_full(0);
    make_psi(0);
    release_and_refill(0);      // column xa-1 only feeds psi
    wait_full(1);
    make_psi(1);
    double psc = psn;
    bool gpc = gpn;
    __syncthreads();

    const int oym = g.wy(y - 1) - y, oyp = g.wy(y + 1) - y;

    for (int r = 1; r <= nplanes; ++r) {
        double fc[9];
        {
            const uint32_t st = stage_a + (r % NS) * C::STAGE_BYTES;
#pragma unroll
            for (int k = 0; k < 9; ++k) fc[k] = lds_f64(st + (k * C::BY + own_box) * 8);
        }
        release_and_refill(r);
        wait_full(r + 1);
        make_psi(r + 1);
        __syncthreads();

        const int sm = (r + 3) & 3, s0 = r & 3, sp = (r + 1) & 3;
        if (inside && ring[s0][tid + 1] >= 0.0) {
            ScForceSums s = {{0., 0., 0.}, {0., 0., 0.}, 0u};
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                if (k == L2::REST) continue;
                const int slot = L2::cx(k) < 0 ? sm : (L2::cx(k) > 0 ? sp : s0);
                const double v = ring[slot][tid + 1 + L2::cy(k)];
                if constexpr (GUO) {
                    const int oslot = L2::cx(k) < 0 ? sp : (L2::cx(k) > 0 ? sm : s0);
                    const double vo = ring[oslot][tid + 1 - L2::cy(k)];
                    if (v < 0.0) s.wall |= 1u << k;
                    sc_force_add<L2>(s, k, false, v < 0.0 ? fmax(vo, 0.0) : v);
                } else {
                    sc_force_add<L2>(s, k, v < 0.0, v);
                }
            }
            double out[9];
            if constexpr (GUO) scrt_collide<L2>(mp, fc, s, Mom<L2>::sum(fc), psc, out);
            else sc_collide<L2>(mp, fc, s, psc, gpc, out);
            const int x = xa - 1 + r;
            const int i = (x + G) * ny + y;
            const unsigned i0 = (unsigned)i, im = (unsigned)(i + (g.wx(x - 1) - x) * ny), ip = (unsigned)(i + (g.wx(x + 1) - x) * ny);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                if (k == L2::REST) { P.base[i0 + P.kn[k]] = out[k]; continue; }
                if (s.wall & (1u << k)) { P.base[i0 + P.kn[L2::opp(k)]] = out[k]; continue; }   // half-way bounce-back
                unsigned idx = (L2::cx(k) < 0 ? im : (L2::cx(k) > 0 ? ip : i0)) + P.kn[k];
                if (L2::cy(k)) idx += (unsigned)(L2::cy(k) < 0 ? oym : oyp);
                P.base[idx] = out[k];
            }
        }
        psc = psn;
        gpc = gpn;
    }
}

bool sc2d_tma_eligible(const clbm_ctx *c)
{
    const Geom &g = c->geo;
    return c->Q == 9 && !c->multi && g.nz == 1 && (g.ny % 2 == 0) && g.ny >= 64 && 9ull * (unsigned long long)g.ncs < (1ull << 32) &&
           c->prm.collision == CLBM_COLLISION_BGK && get_encode() != nullptr;
}

// tensor map of the [9][nx + 2G][ny] population array, its rows described as ny / 2 PAIRS of doubles -- dims (2, ny / 2, nx + 2G, 9) --
// because a box dimension is limited to 256 elements and a column box has T + 4 rows; the bytes land in shared memory in the
// same order.  Cached per buffer like the 3-D lattice's maps (promo = -2 marks this form).
static int cached_tmap_2d(clbm_ctx *c, const void *base, int box_rows, bool pairs, CUtensorMap *out)
{
    const unsigned bx[4] = {pairs ? 2u : (unsigned)box_rows, pairs ? (unsigned)box_rows / 2u : 1u, pairs ? 1u : 9u, pairs ? 9u : 0u};
    const int tag = pairs ? -2 : -3;
    for (const TmapEntry &e : c->tmaps)
        if (e.base == base && e.promo == tag && e.box[0] == bx[0] && e.box[1] == bx[1] && e.box[2] == bx[2] && e.box[3] == bx[3]) {
            memcpy(out, e.map, sizeof(CUtensorMap));
            return 0;
        }
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return CLBM_ECUDA; }
    const Geom &g = c->geo;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r;
    if (pairs) {
        const cuuint64_t dims[4] = {2, (cuuint64_t)g.ny / 2, (cuuint64_t)(g.nx + 2 * g.G), 9};
        const cuuint64_t strides[3] = {16, (cuuint64_t)g.ny * 8, (cuuint64_t)g.ncs * 8};
        const cuuint32_t box[4] = {2, (cuuint32_t)box_rows / 2, 1, 9};
        r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t dims[3] = {(cuuint64_t)g.ny, (cuuint64_t)(g.nx + 2 * g.G), 9};
        const cuuint64_t strides[2] = {(cuuint64_t)g.ny * 8, (cuuint64_t)g.ncs * 8};
        const cuuint32_t box[3] = {(cuuint32_t)box_rows, 1, 9};
        r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (2-D lattice) failed (%d)", (int)r); return CLBM_ECUDA; }
    if (c->tmaps.size() >= 64) c->tmaps.clear();
    TmapEntry e;
    e.base = base;
    for (int i = 0; i < 4; ++i) e.box[i] = bx[i];
    e.promo = tag;
    memcpy(e.map, out, sizeof(CUtensorMap));
    c->tmaps.push_back(e);
    return 0;
}

template <int T, int NS, int MINB, bool GUO>
static int launch_sc2d_tma_g(clbm_ctx *c)
{
    using C = Tma2Cfg<T, NS>;
    const Geom &g = c->geo;
    CUtensorMap tmap;
    if (int rc = cached_tmap_2d(c, c->pop[0][c->parity], C::BY, C::PAIRS, &tmap)) return rc;
    const int tiles = (g.ny + T - 1) / T;
    // short x-chunks keep concurrently resident CTAs on neighbouring columns (their halo rows meet in L2) and give the tail of the
    // grid something to balance with; 32 columns was the optimum of the register-pipelined kernel at 8192^2
    int xchunk = g.nx < 32 ? g.nx : 32;
    if (c->env.sc_xchunk > 0) xchunk = c->env.sc_xchunk < g.nx ? c->env.sc_xchunk : g.nx;
    dim3 grid(tiles, (g.nx + xchunk - 1) / xchunk);
    OutTable2 P = {c->pop[0][1 - c->parity], {0}};
    for (int k = 0; k < 9; ++k) P.kn[k] = (unsigned)((unsigned long long)k * (unsigned long long)g.ncs);
    auto kern = sc2d_tma_kernel<T, NS, MINB, GUO>;
    static PerDeviceOnce attr;
    if (attr.need(c->device)) {
        CLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr.mark(c->device);
    }
    LaunchScope ls(c, "sc2d_tma_collide_stream", true);
    kern<<<grid, T, C::SMEM, c->stream>>>(tmap, P, c->flag, c->pop[0][c->parity], g, c->mp, xchunk);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

template <int T, int NS, int MINB>
static int launch_sc2d_tma(clbm_ctx *c)
{
    if (c->mp.sc_force == CLBM_SC_FORCE_EXPGUO) return launch_sc2d_tma_g<T, NS, MINB, true>(c);
    return launch_sc2d_tma_g<T, NS, MINB, false>(c);
}

// variant: 0 default (128 rows, 3 stages, 6 CTAs per SM: best of tools/sc2d_variants.py); CLBM_SC2D_TMA = 2.. selects other shapes
int sc2d_tma_step(clbm_ctx *c, int variant)
{
    switch (variant) {
    case 2: return launch_sc2d_tma<256, 4, 2>(c);
    case 3: return launch_sc2d_tma<384, 3, 2>(c);
    case 4: return launch_sc2d_tma<384, 2, 2>(c);
    case 5: return launch_sc2d_tma<128, 4, 5>(c);
    case 6: return launch_sc2d_tma<256, 2, 4>(c);
    case 7: return launch_sc2d_tma<256, 3, 3>(c);
    case 8: return launch_sc2d_tma<128, 2, 8>(c);
    case 9: return launch_sc2d_tma<64, 4, 10>(c);
    case 10: return launch_sc2d_tma<256, 3, 4>(c);
    case 11: return launch_sc2d_tma<128, 3, 7>(c);
    case 12: return launch_sc2d_tma<64, 3, 12>(c);
    case 13: return launch_sc2d_tma<224, 3, 3>(c);   // the tallest tiles whose box (228, 196 rows) still fits one TMA dimension
    case 14: return launch_sc2d_tma<192, 3, 4>(c);
    case 15: return launch_sc2d_tma<128, 2, 6>(c);
    default: return launch_sc2d_tma<128, 3, 6>(c);
    }
}

}  // namespace clbm
