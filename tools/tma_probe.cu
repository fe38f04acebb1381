// tma_probe.cu -- stand-alone probe of the 4-D FLOAT64 tensor-map box load used by sc_fused_tma.cu.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int SZ, int SY, int Q>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, double *out, int c0, int c1, int c2)
{
    extern __shared__ __align__(128) unsigned char smem[];
    double *st = reinterpret_cast<double *>(smem);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + ((Q * SY * SZ * 8 + 127) / 128) * 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(Q * SY * SZ * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(smem_u32(st)), "l"(&tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(0) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni WD;\n\tbra.uni WL;\n\tWD:\n\t}" ::"r"(smem_u32(bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < Q * SY * SZ; i += blockDim.x) out[i] = st[i];
}

int main()
{
    const int nz = 36, ny = 24, nxs = 42, Q = 19;
    const int SZ = 34, SY = 10;
    size_t n = (size_t)Q * nxs * ny * nz;
    std::vector<double> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (double)i;
    double *d, *o;
    cudaMalloc(&d, n * 8);
    cudaMalloc(&o, Q * SY * SZ * 8);
    cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    auto enc = (CUresult(*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                            const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))p;
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)nz, (cuuint64_t)ny, (cuuint64_t)nxs, (cuuint64_t)Q};
    cuuint64_t str[3] = {(cuuint64_t)nz * 8, (cuuint64_t)ny * nz * 8, (cuuint64_t)nxs * ny * nz * 8};
    cuuint32_t box[4] = {SZ, SY, 1, Q};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    const int smem = ((Q * SY * SZ * 8 + 127) / 128) * 128 + 64;
    cudaFuncSetAttribute(probe<SZ, SY, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int coords[6][3] = {{2, 0, 3}, {0, -1, 3}, {-2, -1, 3}, {8, 16, 41}, {4, 23, 0}, {-2, 1, 3}};
    for (auto &c : coords) {
        probe<SZ, SY, Q><<<1, 128, smem>>>(tm, o, c[0], c[1], c[2]);
        cudaError_t e = cudaDeviceSynchronize();
        printf("coords (%d,%d,%d): %s\n", c[0], c[1], c[2], cudaGetErrorString(e));
        if (e != cudaSuccess) { printf("  (fault: stopping)\n"); return 1; }
        std::vector<double> ho(Q * SY * SZ);
        cudaMemcpy(ho.data(), o, ho.size() * 8, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int k = 0; k < Q; ++k) for (int sy = 0; sy < SY; ++sy) for (int sz = 0; sz < SZ; ++sz) {
            int z = c[0] + sz, y = c[1] + sy;
            double want = (z < 0 || z >= nz || y < 0 || y >= ny) ? 0.0 : h[(((size_t)k * nxs + c[2]) * ny + y) * nz + z];
            if (ho[(k * SY + sy) * SZ + sz] != want) ++bad;
        }
        printf("  mismatches: %d\n", bad);
    }
    return 0;
}
