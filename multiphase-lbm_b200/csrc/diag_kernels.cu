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

// The scans in GLOBAL x (a slab holds the columns [x_offset, x_offset + nx) of an nx_global-wide lattice; a single slab is the
// special case x_offset = 0, nx = nx_global).  out = {base_y, lstop, rstop, hstop}:
//   base_y  first non-solid row at global x = 0 from y = 2 up (:473-476); ny when this slab does not own x = 0 (combine: MIN).
//           base_y_in >= 0: taken from the caller (second call on a slab, after the MIN over the ranks)
//   lstop   largest global x < xmid on row base_y with rho <= rho_cut, -1 if none here   (combine: MAX)
//   rstop   smallest global x > xmid with rho <= rho_cut, nx_global if none here          (combine: MIN)
//   hstop   first y >= base_y on column xmid that is solid or has rho <= rho_cut; ny when this slab does not own xmid (MIN)
// base = max(0, rstop - lstop - 1), height = hstop - base_y (:489-505); base_y >= ny-1 means "no fluid row" (base = height = 0)
__global__ void __launch_bounds__(DIAG_THREADS)
contact_angle_kernel(const double *__restrict__ fin, const uint8_t *__restrict__ flag, Geom g, double rho_cut, int base_y_in, int *__restrict__ out)
{
    __shared__ int sm[DIAG_THREADS / 32];
    const int nx = g.nx, ny = g.ny, x0 = g.x_offset, xmid = g.nx_global / 2;
    int base_y = base_y_in;
    if (base_y_in < 0) {
        int c = ny;
        if (x0 == 0)
            for (int y = 2 + threadIdx.x; y < ny; y += DIAG_THREADS)
                if (flag[g.idx(0, y, 0)] != CELL_BB) { c = y; break; }
        base_y = block_min(c, sm);
    }
    if (base_y >= ny - 1) {
        if (threadIdx.x == 0) { out[0] = base_y; out[1] = -1; out[2] = g.nx_global; out[3] = ny; }
        return;
    }
    // the walk left / right from xmid stops in front of the first node with rho <= rho_cut (:494-496)
    int lstop = -1, rstop = g.nx_global;
    for (int x = threadIdx.x; x < nx; x += DIAG_THREADS) {
        const int xg = x0 + x;
        if (xg == xmid) continue;
        if (!(node_sum9(fin, g, x, base_y) > rho_cut)) {
            if (xg < xmid) lstop = max(lstop, xg);
            else rstop = min(rstop, xg);
        }
    }
    lstop = block_max(lstop, sm);
    rstop = block_min(rstop, sm);
    // height along xmid: consecutive nodes from base_y that are fluid and denser than rho_cut (:499-505)
    int hstop = ny;
    const int xl = xmid - x0;
    if (xl >= 0 && xl < nx)
        for (int y = base_y + threadIdx.x; y < ny; y += DIAG_THREADS)
            if (flag[g.idx(xl, y, 0)] == CELL_BB || !(node_sum9(fin, g, xl, y) > rho_cut)) { hstop = y; break; }
    hstop = block_min(hstop, sm);
    if (threadIdx.x == 0) { out[0] = base_y; out[1] = lstop; out[2] = rstop; out[3] = hstop; }
}

// blockIdx.x = 0: global column x = 0, 1: global column x = nx_global/2.  out[b] = largest y in [1, ny-2] with phi <= phi_mid,
// 0 if none (:683-706) -- also 0 when this slab does not own the column, so the ranks of a ring combine with MAX
__global__ void __launch_bounds__(DIAG_THREADS)
interface_heights_kernel(const double *__restrict__ fin, Geom g, double phi_mid, int *__restrict__ out)
{
    __shared__ int sm[DIAG_THREADS / 32];
    const int x = (blockIdx.x ? g.nx_global / 2 : 0) - g.x_offset;
    int c = 0;
    if (x >= 0 && x < g.nx)
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

int diag_contact_angle_raw(clbm_ctx *c, double rho_cut, int base_y_in, int out[4])
{
    if (c->prm.model != CLBM_MODEL_SC_D2Q9) { set_error("the base/height contact-angle scan is a Shan-Chen D2Q9 diagnostic"); return CLBM_EINVAL; }
    {
        LaunchScope ls(c, "contact_angle_scan");
        contact_angle_kernel<<<1, DIAG_THREADS, 0, c->stream>>>(c->pop[0][c->parity], c->flag, c->geo, rho_cut, base_y_in, (int *)c->red_dev);
        CLBM_CUDA(cudaGetLastError());
    }
    return fetch_ints(c, out, 4);
}

int diag_contact_angle(clbm_ctx *c, double rho_cut, int *base_y, int *base, int *height)
{
    if (c->multi) { set_error("on an x-slab the contact-angle scan is clbm_diag_contact_angle_slab (two calls, combined by the caller)"); return CLBM_ESTATE; }
    int h[4];
    const int rc = diag_contact_angle_raw(c, rho_cut, -1, h);
    if (rc) return rc;
    const bool none = h[0] >= c->geo.ny - 1;
    if (base_y) *base_y = h[0];
    if (base) *base = none ? 0 : (h[2] - h[1] - 1 > 0 ? h[2] - h[1] - 1 : 0);
    if (height) *height = none ? 0 : h[3] - h[0];
    return 0;
}

int diag_interface_heights(clbm_ctx *c, double phi_mid, int *y_x0, int *y_xmid)
{
    if (c->prm.model != CLBM_MODEL_HCZ_D2Q9) { set_error("the interface-height scan is an HCZ D2Q9 diagnostic"); return CLBM_EINVAL; }
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
