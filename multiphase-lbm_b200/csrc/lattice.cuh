// lattice.cuh -- D2Q9 / D3Q19 descriptors and slab indexing shared by all kernels.
//
// Direction sets, opposites and weights follow the reference tables
//   D2Q9 : SC/apps/laplace2D.h:29-41   (c, opp = k+5 for k<4, rest k=4)
//   D3Q19: PF/apps/laplace3D.h:31-55   (c, opp = k+10 for k<9, rest k=9)
// They are compile-time constants here so that every loop over k unrolls into
// straight-line code with the 0/+-1 factors folded away.
#pragma once
#include <cstddef>
#include <cstdint>

#define CLBM_HD __host__ __device__ __forceinline__
#ifdef CLBM_HOST_CHECK
// tests/host_check only (test infrastructure, never part of libclbm.so): the per-cell device functions are also compiled
// for the host so that their arithmetic can be checked against the oracle on a machine without a GPU
#define CLBM_D __host__ __device__ inline
#else
#define CLBM_D __device__ __forceinline__
#endif

namespace clbm {

struct D2Q9 {
    static constexpr int D = 2, Q = 9, H = 4, REST = 4;
    CLBM_HD static constexpr int cx(int k) { constexpr int v[9] = {-1, 0, -1, -1, 0, 1, 0, 1, 1}; return v[k]; }
    CLBM_HD static constexpr int cy(int k) { constexpr int v[9] = {0, -1, -1, 1, 0, 0, 1, 1, -1}; return v[k]; }
    CLBM_HD static constexpr int cz(int) { return 0; }
    CLBM_HD static constexpr int opp(int k) { constexpr int v[9] = {5, 6, 7, 8, 4, 0, 1, 2, 3}; return v[k]; }
    CLBM_HD static constexpr double t(int k)
    {
        constexpr double v[9] = {1. / 9., 1. / 9., 1. / 36., 1. / 36., 4. / 9., 1. / 9., 1. / 9., 1. / 36., 1. / 36.};
        return v[k];
    }
};

struct D3Q19 {
    static constexpr int D = 3, Q = 19, H = 9, REST = 9;
    CLBM_HD static constexpr int cx(int k)
    {
        constexpr int v[19] = {-1, 0, 0, -1, -1, -1, -1, 0, 0, 0, 1, 0, 0, 1, 1, 1, 1, 0, 0};
        return v[k];
    }
    CLBM_HD static constexpr int cy(int k)
    {
        constexpr int v[19] = {0, -1, 0, -1, 1, 0, 0, -1, -1, 0, 0, 1, 0, 1, -1, 0, 0, 1, 1};
        return v[k];
    }
    CLBM_HD static constexpr int cz(int k)
    {
        constexpr int v[19] = {0, 0, -1, 0, 0, -1, 1, -1, 1, 0, 0, 0, 1, 0, 0, 1, -1, 1, -1};
        return v[k];
    }
    CLBM_HD static constexpr int opp(int k)
    {
        constexpr int v[19] = {10, 11, 12, 13, 14, 15, 16, 17, 18, 9, 0, 1, 2, 3, 4, 5, 6, 7, 8};
        return v[k];
    }
    CLBM_HD static constexpr double t(int k)
    {
        constexpr double v[19] = {1. / 18., 1. / 18., 1. / 18., 1. / 36., 1. / 36., 1. / 36., 1. / 36., 1. / 36., 1. / 36.,
                                  1. / 3.,
                                  1. / 18., 1. / 18., 1. / 18., 1. / 36., 1. / 36., 1. / 36., 1. / 36., 1. / 36., 1. / 36.};
        return v[k];
    }
};

// Slab geometry in device storage.  Storage holds G ghost x-planes on each side of the nx
// local planes: storage plane = x + G, x in [-G, nx+G).  Cell index (z fastest, like the
// reference's i = z + nz*(y + ny*x), PF/apps/laplace3D.h:150-152):
//     s = (x+G)*plane + y*nz + z,   plane = ny*nz.
// wrapx = 1: single slab, periodic in x inside the slab (x+cx wraps into [0,nx)).
// wrapx = 0: x-1 / x+nx live in ghost planes that the halo exchange fills.
// y and z always wrap (walls are mask rows, so bulk nodes never actually wrap through them).
struct Geom {
    int nx, ny, nz, G, wrapx;
    int nx_global, x_offset;
    long long plane;   // ny*nz
    long long ncs;     // (nx+2G)*plane  cells in storage

    CLBM_HD long long idx(int x, int y, int z) const { return (long long)(x + G) * plane + (long long)y * nz + z; }
    CLBM_HD int wx(int x) const
    {
        if (wrapx) { if (x < 0) x += nx; else if (x >= nx) x -= nx; }
        return x;
    }
    CLBM_HD int wy(int y) const { return y < 0 ? y + ny : (y >= ny ? y - ny : y); }
    CLBM_HD int wz(int z) const { return z < 0 ? z + nz : (z >= nz ? z - nz : z); }
    CLBM_HD long long nb(int x, int y, int z, int cx, int cy, int cz) const
    {
        return idx(wx(x + cx), wy(y + cy), wz(z + cz));
    }
};

// c_k . a without multiplications and without the zero components: 0 * a cannot be folded under IEEE rules, so the plain
// c_x a_0 + c_y a_1 + c_z a_2 costs three FP64 operations whatever c_k is; k is a constant after unrolling
template <class L>
CLBM_HD double cdot(int k, double a0, double a1, double a2)
{
    double s = 0.0;
    bool have = false;
    if (L::cx(k) != 0) { s = L::cx(k) > 0 ? a0 : -a0; have = true; }
    if (L::cy(k) != 0) { const double v = L::cy(k) > 0 ? a1 : -a1; s = have ? s + v : v; have = true; }
    if (L::cz(k) != 0) { const double v = L::cz(k) > 0 ? a2 : -a2; s = have ? s + v : v; }
    return s;
}

constexpr uint8_t CELL_BB = 0;    // CellType::bounce_back
constexpr uint8_t CELL_BULK = 1;  // CellType::bulk

}  // namespace clbm
