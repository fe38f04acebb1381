// sc_tma.cuh -- what the TMA-staged D3Q19 Shan-Chen kernels (sc_fused_tma.cu: one CTA per tile and x-chunk; sc_fused_tma_persist.cu:
// persistent CTAs with a work queue) share: the output table, the shared-memory layout, the group barrier.
#pragma once
#include <cuda.h>

#include "sc_cell.cuh"
#include "tma.cuh"

namespace clbm {

using L3 = D3Q19;

// out buffer: direction k starts at k * ncs (one base pointer instead of 19 parameter-block pointers, whose constant-bank loads
// the stores then wait on; the same change in the HCZ D3Q19 sweep kernel halved its long-scoreboard stalls)
struct OutTable {
    double *base;
    size_t ncs;
    unsigned kn[19];   // k * ncs as 32-bit element offsets (valid while 19 * ncs < 2^32: IDX32 kernels)
    CLBM_D double *at(int k) const { return base + (size_t)k * ncs; }
};

template <int TY, int TZ, int NS = 2, int SPY = 1, int SPZ = 1>
struct TmaCfg {
    static constexpr int NT = TY * TZ, SY = TY + 2, SZ = TZ + 2;
    // the innermost TMA coordinate must be 16-byte aligned (odd z faults on sm_100a, tools/tma_probe.cu), so the
    // box spans z0-2 .. z0+TZ+1: BZ = TZ+4 columns, of which column 0 and TZ+3 are padding
    static constexpr int BZ = TZ + 4;
    // SPY x SPZ groups of GY rows x GZ columns, each with its own psi ring (the cells next to a neighbouring group are recomputed)
    static constexpr int SPLIT = SPY * SPZ, GY = TY / SPY, GZ = TZ / SPZ, GT = NT / SPLIT, RY = GY + 2, RZ = GZ + 2;
    static constexpr int NH = 2 * RZ + 2 * GY;                 // halo ring cells of a group
    static constexpr int BOX = 19 * SY * BZ;                   // doubles per staged box
    static_assert(TZ % 2 == 0, "box rows must be a multiple of 16 bytes");
    static_assert(TY % SPY == 0 && TZ % SPZ == 0 && GT % 32 == 0, "groups are whole warps");
    static constexpr int STAGE_BYTES = ((BOX * 8 + 127) / 128) * 128;
    static constexpr int RING_BYTES = SPLIT * 4 * RY * RZ * 8;
    static constexpr int SMEM = NS * STAGE_BYTES + RING_BYTES + 256;   // NS full + NS empty mbarriers + NS counters behind the rings
    static_assert(NH <= GT, "one halo cell per thread");
    static_assert(NS >= 2 && NS <= 8, "2 to 8 stages");
};

// barrier + OR over the threads of one group (named barrier `id`, `nthreads` threads; id 0 = the whole CTA)
template <int SPLIT, int GT>
CLBM_D int group_sync_or(int pred, int id)
{
    if (SPLIT == 1) return __syncthreads_or(pred);
    int res;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.s32 q, %1, 0;\n\t"
        "bar.red.or.pred p, %2, %3, q;\n\t"
        "selp.s32 %0, 1, 0, p;\n\t}"
        : "=r"(res) : "r"(pred), "r"(id), "r"(GT) : "memory");
    return res;
}

}  // namespace clbm
