// diag_kernels.cu -- device-side line scans of the two geometric diagnostics the reference drivers evaluate on the host
// every out_freq iterations, so that a driver no longer downloads a whole field to read three integers:
//   contact angle, base/height method   SC/apps/contactAngle2D.h:465-529 (calculateContactAngle)
//   spike / bubble interface heights    PF/apps/rayleighTaylor2D.h:668-708 (findInterfaceHeights)
// Both are serial scans along one row / column in the reference ("walk until the field drops below the threshold"); here
// each becomes a min / max reduction over the line (first index where the walk would stop), one thread block per line,
// warp-shuffle + shared-memory reduction.  The scanned scalar is rho = phi = sum_k f_k with the reference's association
// (moments.cuh), taken from the current "in" populations, so the integers are exactly those of the reference's scan.
#include <climits>
#include <cstring>

#include "clbm_internal.h"
#include "moments.cuh"

namespace clbm {

constexpr int DIAG_THREADS = 256;

CLBM_D int block_min(int v, int *sm)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();                 // sm may still be read from the previous reduction
    if (l == 0) sm[w] = v;
    __syncthreads();
    v = sm[l < DIAG_THREADS / 32 ? l : 0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;                        // every thread holds the block minimum
}
CLBM_D int block_max(int v, int *sm) { return -block_min(-v, sm); }

CLBM_D double node_sum9(const double *__restrict__ fin, const Geom &g, int x, int y)
{
    double f[9];
    const long long i = g.idx(x, y, 0);
#pragma unroll
    for (int k = 0; k < 9; ++k) f[k] = fin[(size_t)k * g.ncs + i];
    return Mom<D2Q9>::sum(f);
}

// out = {base_y, base, height}; base_y >= ny-1 means "no fluid row found above wall" (base = height = 0 then)
__global__ void __launch_bounds__(DIAG_THREADS)
contact_angle_kernel(const double *__restrict__ fin, const uint8_t *__restrict__ flag, Geom g, double rho_cut, int *__restrict__ out)
{
    __shared__ int sm[DIAG_THREADS / 32];
    const int nx = g.nx, ny = g.ny, xmid = nx / 2;
    // 1) first non-solid row above the bottom wall, looked for at x = 0 from y = 2 upwards (:473-476)
    int c = ny;
    for (int y = 2 + threadIdx.x; y < ny; y += DIAG_THREADS)
        if (flag[g.idx(0, y, 0)] != CELL_BB) { c = y; break; }
    const int base_y = block_min(c, sm);
    if (base_y >= ny - 1) {
        if (threadIdx.x == 0) { out[0] = base_y; out[1] = 0; out[2] = 0; }
        return;
    }
    // 3) the walk left / right from xmid stops in front of the first node with rho <= rho_cut (:494-496)
    int lstop = -1, rstop = nx;     // largest x < xmid / smallest x > xmid that stops the walk
    for (int x = threadIdx.x; x < nx; x += DIAG_THREADS) {
        if (x == xmid) continue;
        if (!(node_sum9(fin, g, x, base_y) > rho_cut)) {
            if (x < xmid) lstop = max(lstop, x);
            else rstop = min(rstop, x);
        }
    }
    lstop = block_max(lstop, sm);
    rstop = block_min(rstop, sm);
    const int left = lstop + 1, right = rstop - 1;
    // 4) height along xmid: consecutive nodes from base_y that are fluid and denser than rho_cut (:499-505)
    int hstop = ny;
    for (int y = base_y + threadIdx.x; y < ny; y += DIAG_THREADS)
        if (flag[g.idx(xmid, y, 0)] == CELL_BB || !(node_sum9(fin, g, xmid, y) > rho_cut)) { hstop = y; break; }
    hstop = block_min(hstop, sm);
    if (threadIdx.x == 0) { out[0] = base_y; out[1] = max(0, right - left + 1); out[2] = hstop - base_y; }
}

// blockIdx.x = 0: column x = 0, 1: column x = nx/2.  out[b] = largest y in [1, ny-2] with phi <= phi_mid, 0 if none (:683-706)
__global__ void __launch_bounds__(DIAG_THREADS)
interface_heights_kernel(const double *__restrict__ fin, Geom g, double phi_mid, int *__restrict__ out)
{
    __shared__ int sm[DIAG_THREADS / 32];
    const int x = blockIdx.x ? g.nx / 2 : 0;
    int c = 0;
    for (int y = 1 + threadIdx.x; y <= g.ny - 2; y += DIAG_THREADS)
        if (node_sum9(fin, g, x, y) <= phi_mid) c = y;     // y grows along the loop: the last hit is this thread's largest
    c = block_max(c, sm);
    if (threadIdx.x == 0) out[blockIdx.x] = c;
}

static int fetch_ints(clbm_ctx *c, int *host, int n)
{
    cudaError_t e = cudaMemcpyAsync(c->red_host, c->red_dev, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "diagnostic scan", __FILE__, __LINE__);
    memcpy(host, c->red_host, n * sizeof(int));
    return 0;
}

int diag_contact_angle(clbm_ctx *c, double rho_cut, int *base_y, int *base, int *height)
{
    if (c->prm.model != CLBM_MODEL_SC_D2Q9) { set_error("the base/height contact-angle scan is a Shan-Chen D2Q9 diagnostic"); return CLBM_EINVAL; }
    if (c->multi) { set_error("the contact-angle scan walks along x: single-slab contexts only"); return CLBM_ESTATE; }
    {
        LaunchScope ls(c, "contact_angle_scan");
        contact_angle_kernel<<<1, DIAG_THREADS, 0, c->stream>>>(c->pop[0][c->parity], c->flag, c->geo, rho_cut, (int *)c->red_dev);
        CLBM_CUDA(cudaGetLastError());
    }
    int h[3];
    const int rc = fetch_ints(c, h, 3);
    if (rc) return rc;
    if (base_y) *base_y = h[0];
    if (base) *base = h[1];
    if (height) *height = h[2];
    return 0;
}

int diag_interface_heights(clbm_ctx *c, double phi_mid, int *y_x0, int *y_xmid)
{
    if (c->prm.model != CLBM_MODEL_HCZ_D2Q9) { set_error("the interface-height scan is an HCZ D2Q9 diagnostic"); return CLBM_EINVAL; }
    if (c->multi) { set_error("the interface-height scan reads the columns x = 0 and x = nx/2 of the whole lattice: single-slab contexts only"); return CLBM_ESTATE; }
    {
        LaunchScope ls(c, "interface_heights_scan");
        interface_heights_kernel<<<2, DIAG_THREADS, 0, c->stream>>>(c->pop[0][c->parity], c->geo, phi_mid, (int *)c->red_dev);
        CLBM_CUDA(cudaGetLastError());
    }
    int h[2];
    const int rc = fetch_ints(c, h, 2);
    if (rc) return rc;
    if (y_x0) *y_x0 = h[0];
    if (y_xmid) *y_xmid = h[1];
    return 0;
}

}  // namespace clbm
