// hcz3d_edges_host.cu -- TEST INFRASTRUCTURE (never part of libclbm.so): the partial-sum scheme of the single-sweep HCZ D3Q19
// kernel (csrc/hcz3d_edges.cuh: gather_pushes, fold_pushes, finish_*, ring_slot, edge_offsets) compiled for the HOST and run
// tile by tile, plane by plane, exactly in the order a CTA of hcz3d_sweep_kernel runs it.  tests/test_hcz3d_edges.py feeds it
// random post-collision populations and compares the assembled moments with a direct periodic gather.
#define CLBM_HOST_CHECK 1
#include <cstring>
#include <vector>

#include "../../multiphase-lbm_b200/csrc/hcz3d_edges.cuh"

using namespace clbm;

template <int TY, int TZ>
static int run(int nx, int ny, int nz, const double *post, double *out, int wrap)
{
    if (ny % TY || nz % TZ || nx < 2) return -1;
    const EdgeGeom eg = make_edge_geom<TY, TZ>(ny, nz);
    const size_t plane = (size_t)ny * nz, nelem = (size_t)nx * plane;
    std::vector<double> M(5 * nelem, 0.0), E((size_t)5 * nx * eg.eplane, 0.0);
    constexpr int NT = TY * TZ, Z1 = TZ + 2, Y1 = TY + 2, NH1 = 2 * Z1 + 2 * TY;
    for (int y0 = 0; y0 < ny; y0 += TY)
        for (int z0 = 0; z0 < nz; z0 += TZ) {
            std::vector<double> T((size_t)5 * (NT + NH1), 0.0), A((size_t)4 * (NT + NH1), 0.0), S((size_t)38 * NT);
            // ring cells in the kernel's order: top row, bottom row, then the two columns
            int rdy[NH1], rdz[NH1];
            for (int t = 0; t < NH1; ++t) {
                int h1y, h1z;
                if (t < Z1) { h1y = 0; h1z = t; }
                else if (t < 2 * Z1) { h1y = Y1 - 1; h1z = t - Z1; }
                else { const int q = t - 2 * Z1; h1y = 1 + (q >> 1); h1z = (q & 1) ? Z1 - 1 : 0; }
                rdy[t] = h1y - 1; rdz[t] = h1z - 1;
            }
            auto cell = [&](int slot, int dy, int dz, int xsrc, bool ring) {
                PushSums s;
                gather_pushes<TY, TZ>(S.data(), dy, dz, s);
                double t5[5], a4[4], v[5];
                for (int j = 0; j < 5; ++j) t5[j] = T[(size_t)j * (NT + NH1) + slot];
                for (int j = 0; j < 4; ++j) a4[j] = A[(size_t)j * (NT + NH1) + slot];
                fold_pushes(t5, a4, s, xsrc == 0, v);
                const int xp = xsrc >= 1 ? xsrc - 1 : (wrap ? nx - 1 : -1);
                if (xp >= 0)
                    for (int m = 0; m < 5; ++m) {
                        if (ring) E[((size_t)m * nx + xp) * eg.eplane + ring_slot<TY, TZ>(eg, y0, z0, dy, dz)] = v[m];
                        else M[(size_t)m * nelem + xp * plane + (size_t)(y0 + dy) * nz + z0 + dz] = v[m];
                    }
                for (int j = 0; j < 5; ++j) T[(size_t)j * (NT + NH1) + slot] = t5[j];
                for (int j = 0; j < 4; ++j) A[(size_t)j * (NT + NH1) + slot] = a4[j];
            };
            for (int xsrc = 0; xsrc < nx; ++xsrc) {
                for (int k = 0; k < 38; ++k)
                    for (int ty = 0; ty < TY; ++ty)
                        for (int tz = 0; tz < TZ; ++tz)
                            S[(size_t)k * NT + ty * TZ + tz] = post[(size_t)k * nelem + xsrc * plane + (size_t)(y0 + ty) * nz + z0 + tz];
                for (int tid = 0; tid < NT; ++tid) cell(tid, tid / TZ, tid % TZ, xsrc, false);
                for (int t = 0; t < NH1; ++t) cell(NT + t, rdy[t], rdz[t], xsrc, true);
            }
            if (wrap) {
                auto fin = [&](int slot, double *arr, size_t stride, size_t i_last, size_t i_first) {
                    double t5[5], a4[4];
                    for (int j = 0; j < 5; ++j) t5[j] = T[(size_t)j * (NT + NH1) + slot];
                    for (int j = 0; j < 4; ++j) a4[j] = A[(size_t)j * (NT + NH1) + slot];
                    for (int m = 0; m < 5; ++m) {
                        arr[m * stride + i_last] = finish_last(t5, m, arr[m * stride + i_last]);
                        arr[m * stride + i_first] = finish_first(a4, m, arr[m * stride + i_first]);
                    }
                };
                for (int tid = 0; tid < NT; ++tid) {
                    const size_t yz = (size_t)(y0 + tid / TZ) * nz + z0 + tid % TZ;
                    fin(tid, M.data(), nelem, (size_t)(nx - 1) * plane + yz, yz);
                }
                for (int t = 0; t < NH1; ++t) {
                    const int rs = ring_slot<TY, TZ>(eg, y0, z0, rdy[t], rdz[t]);
                    fin(NT + t, E.data(), (size_t)nx * eg.eplane, (size_t)(nx - 1) * eg.eplane + rs, (size_t)rs);
                }
            }
        }
    // the consumer's assembly: node array + EY + EZ + EC in that order
    for (int m = 0; m < 5; ++m)
        for (int x = 0; x < nx; ++x)
            for (int y = 0; y < ny; ++y)
                for (int z = 0; z < nz; ++z) {
                    int e[3];
                    edge_offsets<TY, TZ>(eg, y, z, e);
                    double v = M[(size_t)m * nelem + x * plane + (size_t)y * nz + z];
                    const double *Ep = E.data() + ((size_t)m * nx + x) * eg.eplane;
                    const double e0 = e[0] >= 0 ? Ep[e[0]] : 0.0, e1 = e[1] >= 0 ? Ep[e[1]] : 0.0, e2 = e[2] >= 0 ? Ep[e[2]] : 0.0;
                    out[(size_t)m * nelem + x * plane + (size_t)y * nz + z] = ((v + e0) + e1) + e2;
                }
    return 0;
}

extern "C" int host_check_hcz3d_edges(int ty, int tz, int nx, int ny, int nz, const double *post, double *out, int wrap)
{
    if (ty == 8 && tz == 32) return run<8, 32>(nx, ny, nz, post, out, wrap);
    if (ty == 4 && tz == 8) return run<4, 8>(nx, ny, nz, post, out, wrap);
    return -2;
}

// ---- the push of one plane as the kernel does it with VAR bit 2: box stores where push_by_box() says so (emulated with the TMA
//      unit's clipping of out-of-range elements), thread-level stores with the periodic wrap everywhere else.
//      post / out: [19][ny][nz]; writes[k][y][z] counts the stores a slot received.  Returns a non-zero code when a box would
//      break the device's rule for box stores (negative or 16-byte-unaligned start) or would be clipped.
template <int TY, int TZ>
static int run_push(int ny, int nz, const double *post, double *out, int *writes)
{
    if (ny % TY || nz % TZ) return -1;
    const size_t plane = (size_t)ny * nz;
    for (int y0 = 0; y0 < ny; y0 += TY)
        for (int z0 = 0; z0 < nz; z0 += TZ)
            for (int k = 0; k < 19; ++k) {
                const int cy = D3Q19::cy(k), cz = D3Q19::cz(k);
                if (push_by_box<TY, TZ>(y0, ny, cz)) {
                    const PushBox b = push_box_start<TY, TZ>(y0, z0, cy);
                    if (b.y < 0 || b.z < 0) return 1;                       // negative start: traps
                    if ((b.z * 8) % 16) return 2;                           // unaligned start: traps
                    if (b.y + TY > ny || b.z + TZ > nz) return 3;           // would be clipped: elements lost
                    for (int r = 0; r < TY; ++r)
                        for (int c = 0; c < TZ; ++c) {
                            const int yy = b.y + r, zz = b.z + c;
                            if (yy < 0 || yy >= ny || zz < 0 || zz >= nz) continue;
                            out[k * plane + (size_t)yy * nz + zz] = post[k * plane + (size_t)(y0 + r) * nz + (z0 + c)];
                            ++writes[k * plane + (size_t)yy * nz + zz];
                        }
                } else {
                    for (int r = 0; r < TY; ++r)
                        for (int c = 0; c < TZ; ++c) {
                            const int yy = (y0 + r + cy + ny) % ny, zz = (z0 + c + cz + nz) % nz;
                            out[k * plane + (size_t)yy * nz + zz] = post[k * plane + (size_t)(y0 + r) * nz + (z0 + c)];
                            ++writes[k * plane + (size_t)yy * nz + zz];
                        }
                }
            }
    return 0;
}

extern "C" int host_check_hcz3d_box_push(int ty, int tz, int ny, int nz, const double *post, double *out, int *writes)
{
    if (ty == 8 && tz == 32) return run_push<8, 32>(ny, nz, post, out, writes);
    if (ty == 4 && tz == 8) return run_push<4, 8>(ny, nz, post, out, writes);
    return -2;
}

// how many (tile, direction) pushes leave as boxes
extern "C" int host_check_hcz3d_box_count(int ty, int tz, int ny, int nz)
{
    int n = 0;
    for (int y0 = 0; y0 < ny; y0 += ty)
        for (int z0 = 0; z0 < nz; z0 += tz)
            for (int k = 0; k < 19; ++k)
                n += (ty == 8 ? push_by_box<8, 32>(y0, ny, D3Q19::cz(k)) : push_by_box<4, 8>(y0, ny, D3Q19::cz(k))) ? 1 : 0;
    return n;
}
