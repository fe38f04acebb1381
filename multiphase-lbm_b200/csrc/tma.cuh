// tma.cuh -- sm_100a bulk-tensor (TMA) and mbarrier primitives shared by the plane-marching kernels:
// cp.async.bulk.tensor loads of population boxes into shared memory, completion on an mbarrier
// (expect_tx / complete_tx), and the driver entry point that encodes the tensor maps.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include <cstring>

#include "clbm_internal.h"
#include "lattice.cuh"

namespace clbm {

CLBM_D uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

CLBM_D void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
CLBM_D void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
CLBM_D void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
CLBM_D void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
CLBM_D double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
CLBM_D void tma_load_4d(uint32_t dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

CLBM_D void tma_load_3d(uint32_t dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

CLBM_D void tma_load_2d(uint32_t dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// shared -> global box store (bulk async-group of the issuing thread) and its bookkeeping
CLBM_D void tma_store_4d(const CUtensorMap *tmap, uint32_t src, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(tmap), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
CLBM_D void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
CLBM_D void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // the sources may be overwritten
CLBM_D void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
CLBM_D void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }  // generic-proxy smem writes -> async proxy

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// tensor map of the [Q][nx+2G][ny][nz] fp64 population array at `base` with the given box, encoded once per
// (buffer, box, promotion) and kept in the context (the two parities of a population set alternate between two entries)
inline int cached_tmap(clbm_ctx *c, const void *base, const cuuint32_t box[4], CUtensorMapL2promotion promo, CUtensorMap *out)
{
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
    for (const TmapEntry &e : c->tmaps)
        if (e.base == base && e.promo == (int)promo && e.box[0] == box[0] && e.box[1] == box[1] && e.box[2] == box[2] && e.box[3] == box[3]) {
            memcpy(out, e.map, sizeof(CUtensorMap));
            return 0;
        }
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return CLBM_ECUDA; }
    const Geom &g = c->geo;
    const cuuint64_t dims[4] = {(cuuint64_t)g.nz, (cuuint64_t)g.ny, (cuuint64_t)(g.nx + 2 * g.G), (cuuint64_t)c->Q};
    const cuuint64_t strides[3] = {(cuuint64_t)g.nz * 8, (cuuint64_t)g.plane * 8, (cuuint64_t)g.ncs * 8};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return CLBM_ECUDA; }
    if (c->tmaps.size() >= 64) c->tmaps.clear();
    TmapEntry e;
    e.base = base;
    for (int i = 0; i < 4; ++i) e.box[i] = box[i];
    e.promo = (int)promo;
    memcpy(e.map, out, sizeof(CUtensorMap));
    c->tmaps.push_back(e);
    return 0;
}

}  // namespace clbm
