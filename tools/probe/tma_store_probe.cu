// standalone probe: does a cp.async.bulk.tensor shared -> global box store of fp64 data work on this device, with the box and
// the descriptor settings hcz3d_sweep.cu uses (tile mode, no swizzle, box {32, 8, 1, 1}), also with a box that starts at -1?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_store_probe tma_store_probe.cu -lcuda && ./tma_store_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int NZ = 64, NY = 16, NX = 4, Q = 3, TZ = 32, TY = 8;
__device__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, int mode, int cz, int cy, int flags)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + TZ * TY * 8);
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (mode == 0) {   // TMA load, generic read-modify-write, TMA store
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(TZ * TY * 8) : "memory");
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(s32(smem)), "l"(&tin), "r"(s32(bar)), "r"(0), "r"(0), "r"(1), "r"(0) : "memory");
        }
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra.uni D;\n\tbra.uni W;\n\tD:\n\t}" ::"r"(s32(bar)) : "memory");
    }
    double *s = reinterpret_cast<double *>(smem);
    s[tid] = (mode == 0 ? s[tid] : 0.0) + 1000.0 + tid;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == ((flags & 2) ? 0 : blockDim.x - 32)) {
        const int c0 = (flags & 4) ? 2 * cz : cz;     // 32-bit element view: twice the inner coordinate
        if (flags & 1)
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(&tout), "r"(s32(smem)), "r"(c0), "r"(cy), "r"(2), "r"(1) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(&tout), "r"(s32(smem)), "r"(c0), "r"(cy), "r"(2), "r"(1) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                        const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv)
{
    const int flags = argc > 1 ? atoi(argv[1]) : 0, only_mode = argc > 2 ? atoi(argv[2]) : -1, czarg = argc > 3 ? atoi(argv[3]) : 1;
    printf("flags %d (1 = no .tile, 2 = tid 0 issues, 4 = output map as UINT32 elements, 8 = UINT64)\n", flags);
    const size_t n = (size_t)Q * NX * NY * NZ;
    std::vector<double> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (double)i;
    double *din, *dout;
    CK(cudaMalloc(&din, n * 8)); CK(cudaMalloc(&dout, n * 8));
    CK(cudaMemcpy(din, h.data(), n * 8, cudaMemcpyHostToDevice));
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    Enc enc = (Enc)p;
    const cuuint64_t dims[4] = {NZ, NY, NX, Q};
    const cuuint64_t str[3] = {NZ * 8, NZ * NY * 8, (cuuint64_t)NZ * NY * NX * 8};
    const cuuint32_t box[4] = {TZ, TY, 1, 1}, es[4] = {1, 1, 1, 1};
    CUtensorMap tin, tout;
    for (int promo = 0; promo < 2; ++promo) {
        CUresult r1 = enc(&tin, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, din, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const cuuint64_t dims32[4] = {2 * NZ, NY, NX, Q};
        const cuuint32_t box32[4] = {2 * TZ, TY, 1, 1};
        CUresult r2 = enc(&tout, (flags & 4) ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : ((flags & 8) ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64), 4, dout,
                          (flags & 4) ? dims32 : dims, str, (flags & 4) ? box32 : box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d %d (out promo %d)\n", (int)r1, (int)r2, promo);
        for (int mode = 0; mode < 2; ++mode) {
            if (only_mode >= 0 && mode != only_mode) continue;
            for (int neg = 0; neg < 2; ++neg) {
                CK(cudaMemset(dout, 0, n * 8));
                probe<<<1, 256, TZ * TY * 8 + 64>>>(tin, tout, mode, neg ? -czarg : czarg, neg ? -1 : 1, flags);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d neg %d: %s\n", mode, neg, cudaGetErrorString(e)); return 2; }
                std::vector<double> o(n);
                CK(cudaMemcpy(o.data(), dout, n * 8, cudaMemcpyDeviceToHost));
                long bad = 0, written = 0;
                const int c0 = neg ? -1 : 1, cz0 = neg ? -czarg : czarg;
                for (int k = 0; k < Q; ++k) for (int x = 0; x < NX; ++x) for (int y = 0; y < NY; ++y) for (int z = 0; z < NZ; ++z) {
                    const size_t i = ((size_t)(k * NX + x) * NY + y) * NZ + z;
                    double want = 0.0;
                    const int ty = y - c0, tz = z - cz0;
                    if (k == 1 && x == 2 && ty >= 0 && ty < TY && tz >= 0 && tz < TZ) {
                        const int t = ty * TZ + tz;
                        want = 1000.0 + t + (mode == 0 ? h[((size_t)(0 * NX + 1) * NY + ty) * NZ + tz] : 0.0);
                        ++written;
                    }
                    if (o[i] != want) ++bad;
                }
                printf("mode %d (%s) box at (z %d, y %d): %ld elements expected, %ld mismatches\n", mode, mode ? "generic fill" : "TMA load + add", cz0, c0, written, bad);
            }
        }
    }
    return 0;
}
