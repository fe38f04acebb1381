// support_kernels.cu -- everything around the step kernels: device-side initial conditions
// (iniLattice + inigeom of each case), warp-shuffle reductions for the mass / energy diagnostics,
// and the ghost-plane pack / unpack of the x-slab decomposition.
#include "clbm_internal.h"
#include "ring_sync.cuh"
#include "moments.cuh"

namespace clbm {

RingSync ring_sync_for(const clbm_ctx *c, int phase, int mode, unsigned nblocks);   // slab_comm.cu

// ============================================================================================
// initial conditions.  One thread per storage cell INCLUDING ghost planes so that ghost flags are
// consistent with the global geometry from the start (global x = (x + x_offset) mod nx_global).
// ============================================================================================
struct CaseArgs { double a[8]; int n; };

template <class L>
__global__ void __launch_bounds__(256)
init_case_kernel(double *__restrict__ f, double *__restrict__ gpop, double *__restrict__ f1, double *__restrict__ g1,
                 uint8_t *__restrict__ flag, Geom g, ModelParams mp, int case_id, CaseArgs A)
{
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.ncs) return;
    const int xs = (int)(s / g.plane);
    const int r = (int)(s % g.plane);
    const int iY = r / g.nz, iZ = r % g.nz;
    int iX = xs - g.G + g.x_offset;
    iX %= g.nx_global;
    if (iX < 0) iX += g.nx_global;
    const int nx = g.nx_global, ny = g.ny, nz = g.nz;

    double vf = 0.0, vg = 0.0;   // f_k = vf * t_k, g_k = vg * t_k
    bool wall = false, both = false;
    switch (case_id) {
    case CLBM_CASE_SC_LAPLACE2D: {   // SC/apps/laplace2D.h:132-145, 397-404
        const double cx = double(nx) / 2.0, cy = double(ny) / 2.0, Rdrop = A.a[2];
        const double dx = double(iX) - cx, dy = double(iY) - cy;
        vf = (dx * dx + dy * dy <= Rdrop * Rdrop) ? A.a[0] : A.a[1];
    } break;
    case CLBM_CASE_SC_CONTACT2D: {   // SC/apps/contactAngle2D.h:126-137, 442-455
        const int x_c = nx / 2, y_c = 5;
        const double dx = double(iX) - double(x_c), dy = double(iY) - double(y_c);
        vf = (dx * dx + dy * dy <= A.a[2] * A.a[2]) ? A.a[0] : A.a[1];
        wall = (iY == 0 || iY == ny - 1);
    } break;
    case CLBM_CASE_SC_LAYERED2D: {   // SC/apps/twoLayeredFlow2D.h:325-346, 441-454   args {rhol, rhog, h_lower, w_int}
        const double H = double(ny - 1);
        const double hl = A.a[2] < 0.0 ? 0.0 : (A.a[2] > 0.5 ? 0.5 : A.a[2]);
        const double y_low = hl * H, y_high = H - y_low;
        const int wi = (int)A.a[3];
        const double w = double(wi > 1 ? wi : 1), yy = double(iY);
        const double s_bottom = 0.5 * (1.0 - tanh((yy - y_low) / w));
        const double s_top = 0.5 * (1.0 + tanh((yy - y_high) / w));
        double s_liq = s_bottom + s_top;
        s_liq = s_liq < 0.0 ? 0.0 : (s_liq > 1.0 ? 1.0 : s_liq);
        vf = s_liq * A.a[1] + (1.0 - s_liq) * A.a[0];   // "liquid" (rhog in the reference's naming) at the walls
        wall = (iY == 0 || iY == ny - 1);
    } break;
    case CLBM_CASE_SC_RT2D: {   // SC/apps/RayleighTaylor2D.h:134-158, 526-541   args {rhol, rhog}
        const double x = double(iX);
        const double itf = (double(ny) / 2.0) + double(nx) * 0.1 * cos(2.0 * 3.14159265358979323846 * x / double(nx - 1));
        const double w = 2.5, y = double(iY);
        vf = 0.5 * (A.a[0] + A.a[1]) + 0.5 * (A.a[0] - A.a[1]) * tanh((y - itf) / (2.0 * w));
        wall = (iY == 0 || iY == ny - 1);
    } break;
    case CLBM_CASE_SC_DROPLET3D:
    case CLBM_CASE_SC_DROPLET3D_PER: {   // contactAngle2D geometry extruded to 3-D
        const bool per = case_id == CLBM_CASE_SC_DROPLET3D_PER;
        const double yc = per ? double(ny / 2) : (A.n > 3 ? A.a[3] : 5.0);
        const double dx = double(iX) - double(nx / 2), dy = double(iY) - yc, dz = double(iZ) - double(nz / 2);
        vf = (dx * dx + dy * dy + dz * dz <= A.a[2] * A.a[2]) ? A.a[0] : A.a[1];
        wall = !per && (iY == 0 || iY == ny - 1);
    } break;
    case CLBM_CASE_HCZ_RT2D: {   // PF/apps/rayleighTaylor2D.h:155-193, 802-820
        const double x = double(iX);
        const double itf = (double(ny) / 2.0) + double(nx) * 0.1 * cos(2.0 * 3.14159265358979323846 * x / double(nx - 1));
        const double w = 1.25, y = double(iY);
        const double phi = 0.5 * (mp.phi_l + mp.phi_g) + 0.5 * (mp.phi_l - mp.phi_g) * tanh((y - itf) / (2.0 * w));
        const double rho = mp.rho_g + ((phi - mp.phi_g) / (mp.phi_l - mp.phi_g)) * (mp.rho_l - mp.rho_g);
        const double rt = mp.b * rho / 4.0, d = 1.0 - rt;
        vf = phi;
        vg = (rho / 3.0) * (1.0 + rt + rt * rt - rt * rt * rt) / (d * d * d) - mp.a * rho * rho;
        wall = (iY == 0 || iY == ny - 1);
    } break;
    case CLBM_CASE_HCZ_LAYERED2D: {   // PF/apps/twoLayeredFlow2D.h:148-196, 737-757   args {h_lower, w_int}
        const double H = double(ny - 1);
        const double hl = A.a[0] < 0.0 ? 0.0 : (A.a[0] > 0.5 ? 0.5 : A.a[0]);
        const double y_low = hl * H, y_high = H - y_low;
        const int wi = (int)A.a[1];
        const double w = double(wi > 1 ? wi : 1), yy = double(iY);
        double s_liq = 0.5 * (1.0 - tanh((yy - y_low) / w)) + 0.5 * (1.0 + tanh((yy - y_high) / w));
        s_liq = s_liq < 0.0 ? 0.0 : (s_liq > 1.0 ? 1.0 : s_liq);
        const double s_gas = 1.0 - s_liq;
        const double rho = s_liq * mp.rho_g + s_gas * mp.rho_l;
        const double rt = mp.b * rho / 4.0, d = 1.0 - rt;
        vf = s_liq * mp.phi_g + s_gas * mp.phi_l;
        vg = (rho / 3.0) * (1.0 + rt + rt * rt - rt * rt * rt) / (d * d * d) - mp.a * rho * rho;
        wall = (iY == 0 || iY == ny - 1);
        both = true;    // the reference fills fout / gout too (:189-190) and inigeom zeroes only the parity-0 buffer of walls
    } break;
    case CLBM_CASE_HCZ_LAPLACE3D: {   // PF/apps/laplace3D.h:170-213
        const double xc = double(nx) / 2.0, yc = double(ny) / 2.0, zc = double(nz) / 2.0, R = 0.25 * nx;
        const double dx = double(iX) - xc, dy = double(iY) - yc, dz = double(iZ) - zc;
        const double delta = sqrt(dx * dx + dy * dy + dz * dz) - R;
        const double rl = mp.b * mp.phi_l / 4.0, dl = 1.0 - rl;
        const double pth_l = (mp.phi_l / 3.0) * (1 + rl + rl * rl - rl * rl * rl) / (dl * dl * dl) - mp.a * mp.phi_l * mp.phi_l;
        const double rg = mp.b * mp.phi_g / 4.0, dg = 1.0 - rg;
        const double pth_g = (mp.phi_g / 3.0) * (1 + rg + rg * rg - rg * rg * rg) / (dg * dg * dg) - mp.a * mp.phi_g * mp.phi_g;
        const double w = 0.5 * (1.0 - tanh(delta / 1.0));
        vf = mp.phi_g + w * (mp.phi_l - mp.phi_g);
        vg = pth_g + w * (pth_l - pth_g);
    } break;
    default: break;
    }
    flag[s] = wall ? CELL_BB : CELL_BULK;
#pragma unroll
    for (int k = 0; k < L::Q; ++k) {
        f[(size_t)k * g.ncs + s] = wall ? 0.0 : vf * L::t(k);
        if (gpop) gpop[(size_t)k * g.ncs + s] = wall ? 0.0 : vg * L::t(k);
        if (both && f1) f1[(size_t)k * g.ncs + s] = vf * L::t(k);
        if (both && g1) g1[(size_t)k * g.ncs + s] = vg * L::t(k);
    }
}

int model_init_case(clbm_ctx *c, int case_id, const double *args, int nargs)
{
    const int m = c->prm.model;
    const bool sc2 = m == CLBM_MODEL_SC_D2Q9, sc3 = m == CLBM_MODEL_SC_D3Q19;
    const bool ok = (sc2 && (case_id == CLBM_CASE_SC_LAPLACE2D || case_id == CLBM_CASE_SC_CONTACT2D || case_id == CLBM_CASE_SC_LAYERED2D ||
                            case_id == CLBM_CASE_SC_RT2D)) ||
                    (sc3 && (case_id == CLBM_CASE_SC_DROPLET3D || case_id == CLBM_CASE_SC_DROPLET3D_PER)) ||
                    (m == CLBM_MODEL_HCZ_D2Q9 && (case_id == CLBM_CASE_HCZ_RT2D || case_id == CLBM_CASE_HCZ_LAYERED2D)) ||
                    (m == CLBM_MODEL_HCZ_D3Q19 && case_id == CLBM_CASE_HCZ_LAPLACE3D);
    if (!ok) { set_error("case %d does not belong to model %d", case_id, m); return CLBM_EINVAL; }
    if (case_id == CLBM_CASE_SC_RT2D && nargs < 2) { set_error("the Shan-Chen Rayleigh-Taylor case needs {rhol, rhog}"); return CLBM_EINVAL; }
    if ((sc2 || sc3) && case_id != CLBM_CASE_SC_RT2D && nargs < 3) { set_error("Shan-Chen droplet cases need {rhol, rhog, R}"); return CLBM_EINVAL; }
    CaseArgs A;
    A.n = nargs < 8 ? nargs : 8;
    for (int i = 0; i < 8; ++i) A.a[i] = (args && i < A.n) ? args[i] : 0.0;
    const Geom &g = c->geo;
    c->parity = 0;
    for (int s = 0; s < c->sets; ++s) CLBM_CUDA(cudaMemsetAsync(c->pop[s][1], 0, (size_t)c->Q * g.ncs * sizeof(double), c->stream));
    LaunchScope ls(c, "init_case");
    if (c->Q == 9)
        init_case_kernel<D2Q9><<<grid_for(g.ncs, 256), 256, 0, c->stream>>>(c->pop[0][0], c->pop[1][0], c->pop[0][1], c->pop[1][1], c->flag, g, c->mp, case_id, A);
    else
        init_case_kernel<D3Q19><<<grid_for(g.ncs, 256), 256, 0, c->stream>>>(c->pop[0][0], c->pop[1][0], c->pop[0][1], c->pop[1][1], c->flag, g, c->mp, case_id, A);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

// ============================================================================================
// diagnostics: deterministic two-stage reductions (warp shuffle -> block -> one final block)
// ============================================================================================
CLBM_D double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
CLBM_D double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

constexpr int RED_BLOCKS = 1184;   // 8 x 148 SMs
constexpr int RED_THREADS = 256;

// partial[b] = {mass, energy, umax} of block b
__global__ void __launch_bounds__(RED_THREADS)
reduce_stage1(const double *__restrict__ s0, const double *__restrict__ ux, const double *__restrict__ uy,
              const double *__restrict__ uz, const uint8_t *__restrict__ flag, long long n, double *__restrict__ partial)
{
    double m = 0.0, e = 0.0, um = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint8_t fl = flag[i];
        if (fl != CELL_BB) m += s0[i];
        if (fl == CELL_BULK) {
            const double a = ux[i], b = uy[i], cc = uz[i];
            const double q = a * a + b * b + cc * cc;
            e += q;
            um = fmax(um, q);
        }
    }
    __shared__ double sm[3][RED_THREADS / 32];
    m = warp_sum(m); e = warp_sum(e); um = warp_max(um);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sm[0][w] = m; sm[1][w] = e; sm[2][w] = um; }
    __syncthreads();
    if (w == 0) {
        m = l < RED_THREADS / 32 ? sm[0][l] : 0.0;
        e = l < RED_THREADS / 32 ? sm[1][l] : 0.0;
        um = l < RED_THREADS / 32 ? sm[2][l] : 0.0;
        m = warp_sum(m); e = warp_sum(e); um = warp_max(um);
        if (l == 0) { partial[3 * blockIdx.x + 0] = m; partial[3 * blockIdx.x + 1] = e; partial[3 * blockIdx.x + 2] = um; }
    }
}

__global__ void __launch_bounds__(RED_THREADS)
reduce_stage2(const double *__restrict__ partial, int nb, double *__restrict__ out)
{
    double m = 0.0, e = 0.0, um = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        m += partial[3 * i];
        e += partial[3 * i + 1];
        um = fmax(um, partial[3 * i + 2]);
    }
    __shared__ double sm[3][RED_THREADS / 32];
    m = warp_sum(m); e = warp_sum(e); um = warp_max(um);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sm[0][w] = m; sm[1][w] = e; sm[2][w] = um; }
    __syncthreads();
    if (w == 0) {
        m = l < RED_THREADS / 32 ? sm[0][l] : 0.0;
        e = l < RED_THREADS / 32 ? sm[1][l] : 0.0;
        um = l < RED_THREADS / 32 ? sm[2][l] : 0.0;
        m = warp_sum(m); e = warp_sum(e); um = warp_max(um);
        if (l == 0) { out[0] = m; out[1] = e; out[2] = um; }
    }
}

__global__ void compact_flag_kernel(const uint8_t *__restrict__ flag, uint8_t *__restrict__ out, long long off, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = flag[off + i];
}

int model_reduce(clbm_ctx *c, int kind, double *out)
{
    if (kind < CLBM_REDUCE_MASS || kind > CLBM_REDUCE_UMAX) { set_error("bad reduction kind %d", kind); return CLBM_EINVAL; }
    const Geom &g = c->geo;
    const long long n = (long long)g.nx * g.plane;
    // four field arrays + the per-block partial sums, in the context's persistent scratch (no allocation per call)
    double *tmp = nullptr;
    int rc = field_scratch(c, ((size_t)4 * n + (size_t)3 * RED_BLOCKS) * sizeof(double), &tmp);
    if (rc) return rc;
    double *partial = tmp + (size_t)4 * n;
    rc = model_fields(c, tmp, nullptr, nullptr, tmp + n, tmp + 2 * n, tmp + 3 * n);
    if (!rc) {
        {
            LaunchScope ls(c, "reduce_stage1");
            reduce_stage1<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(tmp, tmp + n, tmp + 2 * n, tmp + 3 * n, c->flag + (size_t)g.G * g.plane, n, partial);
        }
        {
            LaunchScope ls(c, "reduce_stage2");
            reduce_stage2<<<1, RED_THREADS, 0, c->stream>>>(partial, RED_BLOCKS, c->red_dev);
        }
        cudaError_t e = cudaMemcpyAsync(c->red_host, c->red_dev, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "reduce", __FILE__, __LINE__);
    }
    if (rc) return rc;
    const double nglob = (double)g.nx_global * (double)g.ny * (double)g.nz;
    if (kind == CLBM_REDUCE_MASS) *out = c->red_host[0];
    else if (kind == CLBM_REDUCE_ENERGY) *out = 0.5 * c->red_host[1] / nglob;   // this slab's share of 0.5*sum(u.u)/nelem
    else *out = sqrt(c->red_host[2]);
    return 0;
}

// ============================================================================================
// x-slab ghost exchange.  side 0 = towards the x-1 neighbour, side 1 = towards x+1.
//   phase 0 : moment halos before the collide  (SC: psi depth 1; HCZ2D: phi depth 2;
//             HCZ3D: phi depth 3 + P_term and the raw momentum depth 1)
//   phase 1 : populations that crossed the slab face into a ghost plane during the push
//   phase 2 : node mask, depth G, once after upload
// Planes are contiguous in storage (x slowest), so pack is a handful of D2D copies.
// ============================================================================================
static int n_cross(const clbm_ctx *c) { return c->Q == 9 ? 3 : 5; }
static void cross_dirs(const clbm_ctx *c, int side, int *ks)
{
    // directions with c_x = -1 (side 0) / +1 (side 1)
    if (c->Q == 9) { const int m[3] = {0, 2, 3}, p[3] = {5, 7, 8}; for (int i = 0; i < 3; ++i) ks[i] = side ? p[i] : m[i]; }
    else { const int m[5] = {0, 3, 4, 5, 6}, p[5] = {10, 13, 14, 15, 16}; for (int i = 0; i < 5; ++i) ks[i] = side ? p[i] : m[i]; }
}

struct HaloField { int fld, depth; };
static int phase0_fields(const clbm_ctx *c, HaloField *hf)
{
    switch (c->prm.model) {
    case CLBM_MODEL_SC_D2Q9:
    case CLBM_MODEL_SC_D3Q19: hf[0] = {0, 1}; return 1;
    case CLBM_MODEL_HCZ_D2Q9: hf[0] = {0, 2}; return 1;
    case CLBM_MODEL_HCZ_D3Q19: hf[0] = {0, 3}; hf[1] = {1, 1}; hf[2] = {2, 1}; hf[3] = {3, 1}; hf[4] = {4, 1}; return 5;
    }
    return 0;
}

// All halo buffers of a context live in ONE allocation, the "mailbox": [phase][side][send, recv] blocks (256-byte aligned)
// followed by a page of flag words.  One allocation = one cudaIpcMemHandle, and the block offsets depend only on the plane size
// and the model, so a ring neighbour can address our receive blocks and flags through its mapping of the mailbox
// (slab_comm.cu: the peer-memory ring packs straight into the neighbour's receive block).
size_t halo_block_offset(const clbm_ctx *c, int phase, int side, int recv)
{
    size_t off = 0;
    for (int ph = 0; ph < 3; ++ph)
        for (int sd = 0; sd < 2; ++sd)
            for (int r = 0; r < 2; ++r) {
                if (ph == phase && sd == side && r == recv) return off;
                off += (c->halo_bytes[ph] + 255) / 256 * 256;
            }
    return off;   // (3, 0, 0): the flag page
}

int halo_alloc(clbm_ctx *c)
{
    const Geom &g = c->geo;
    HaloField hf[8];
    const int nf = phase0_fields(c, hf);
    size_t planes0 = 0;
    for (int i = 0; i < nf; ++i) planes0 += hf[i].depth;
    c->halo_bytes[0] = planes0 * g.plane * sizeof(double);
    c->halo_bytes[1] = (size_t)n_cross(c) * c->sets * g.plane * sizeof(double);
    c->halo_bytes[2] = (size_t)g.G * g.plane;
    c->mailbox_flags_off = halo_block_offset(c, 3, 0, 0);
    c->mailbox_bytes = c->mailbox_flags_off + 4096;
    if (c->fld0_in_mailbox) {   // Shan-Chen: the psi field behind the flag page (see clbm_internal.h)
        c->mailbox_psi_off = c->mailbox_bytes;
        c->mailbox_bytes += (size_t)g.ncs * sizeof(double);
    }
    if (cudaMalloc(&c->mailbox, c->mailbox_bytes) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory (halo buffers)"); return CLBM_ENOMEM; }
    cudaMemsetAsync(c->mailbox, 0, c->mailbox_bytes, c->stream);
    if (c->fld0_in_mailbox) c->fld[0] = (double *)((char *)c->mailbox + c->mailbox_psi_off);
    for (int ph = 0; ph < 3; ++ph)
        for (int side = 0; side < 2; ++side)
            for (int r = 0; r < 2; ++r) c->halo[ph][side][r] = (char *)c->mailbox + halo_block_offset(c, ph, side, r);
    return 0;
}

// where pack writes what travels towards `side`: our own send block, or -- on a peer-memory ring -- the receive block of
// that neighbour (its block of the OPPOSITE side) through the mapping of its mailbox
void *halo_send_ptr(const clbm_ctx *c, int phase, int side)
{
    if (phase == 0 && c->halo0_direct) {
        // Shan-Chen on a peer ring: our boundary psi plane goes straight into the neighbour's ghost plane -- its RIGHT ghost
        // plane (storage plane nx + G of ITS slab) for what we send to the left, its left ghost plane (storage plane G - 1) to the right
        const size_t pl = (size_t)c->geo.plane * sizeof(double);
        const size_t plane_idx = side == 0 ? (size_t)(c->peer_nx[0] + c->geo.G) : (size_t)(c->geo.G - 1);
        return (char *)c->peer_base[side] + c->mailbox_psi_off + plane_idx * pl;
    }
    if (c->peer_mode && c->peer_base[side]) return (char *)c->peer_base[side] + halo_block_offset(c, phase, 1 - side, 1);
    return c->halo[phase][side][0];
}

// The populations that cross a slab face: slot = set * ncross + i, direction ks[side][i] travels towards `side`.
struct CrossTable {
    double *pop[2];              // population set s of the current "in" buffer
    const double *recv[2];       // receive buffer of side 0 / 1
    double *send[2];             // send buffer of side 0 / 1
    int ks[2][5];                // crossing directions sent towards side 0 / 1 (ks[1 - side] arrive from side)
    int ncross, sets;
};

// directions with c_x = -1 (travel towards side 0) / +1 (towards side 1), as compile-time lists
template <class L> struct CrossDirs;
template <> struct CrossDirs<D2Q9> {
    static constexpr int NC = 3;
    CLBM_HD static constexpr int k(int side, int i) { constexpr int m[3] = {0, 2, 3}, p[3] = {5, 7, 8}; return side ? p[i] : m[i]; }
};
template <> struct CrossDirs<D3Q19> {
    static constexpr int NC = 5;
    CLBM_HD static constexpr int k(int side, int i) { constexpr int m[5] = {0, 3, 4, 5, 6}, p[5] = {10, 13, 14, 15, 16}; return side ? p[i] : m[i]; }
};

// One thread per boundary node and side takes ALL slots of the node: its own mask once, the masks of the NC upstream nodes,
// then NC x sets independent loads in flight.  (The first form -- one thread per node and slot, three dependent loads each,
// 10 240 blocks -- needed 113 us for the 21 MB of a 512 x 512 Shan-Chen face, 218 us for HCZ D3Q19: a tenth of a 64-plane step.)
template <class L, int SIDE>
CLBM_D void unpack_cross_node(const CrossTable &T, const uint8_t *__restrict__ flag, const Geom &g, long long r)
{
    constexpr int NC = CrossDirs<L>::NC;
    const int y = (int)(r / g.nz), z = (int)(r % g.nz);
    const int xb = SIDE ? g.nx - 1 : 0;   // boundary plane that receives
    const long long j = g.idx(xb, y, z);
    const uint8_t fj = flag[j];
    uint8_t fs[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int k = CrossDirs<L>::k(1 - SIDE, i);   // arrives from SIDE: moves away from it
        fs[i] = flag[g.idx(xb - L::cx(k), g.wy(y - L::cy(k)), g.wz(z - L::cz(k)))];
    }
    if (fj == CELL_BB) return;
    double v[2][NC];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int i = 0; i < NC; ++i)
            if (s < T.sets) v[s][i] = T.recv[SIDE][(size_t)(s * NC + i) * g.plane + r];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int i = 0; i < NC; ++i)
            if (s < T.sets && fs[i] != CELL_BB) T.pop[s][(size_t)CrossDirs<L>::k(1 - SIDE, i) * g.ncs + j] = v[s][i];
}

template <class L>
__global__ void __launch_bounds__(256)
unpack_cross_kernel(CrossTable T, const uint8_t *__restrict__ flag, Geom g, RingSync rs)
{
    ring_kernel_begin(rs);   // fused ring: the neighbours' crossing populations have arrived
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < g.plane; r += stride) {
        if (blockIdx.y == 0) unpack_cross_node<L, 0>(T, flag, g, r);
        else unpack_cross_node<L, 1>(T, flag, g, r);
    }
    ring_kernel_end(rs);
}

static CrossTable cross_table(clbm_ctx *c)
{
    CrossTable T;
    T.ncross = n_cross(c);
    T.sets = c->sets;
    for (int s = 0; s < 2; ++s) T.pop[s] = c->pop[s < c->sets ? s : 0][c->parity];
    for (int side = 0; side < 2; ++side) {
        T.recv[side] = (const double *)c->halo[1][side][1];
        T.send[side] = (double *)halo_send_ptr(c, 1, side);
        cross_dirs(c, side, T.ks[side]);
    }
    return T;
}

// Up to 24 (dst, src, bytes) segments copied -- or, src == nullptr, zero-filled -- by ONE launch (blockIdx.y = segment).  The
// moment halo of a slab step is 2 (Shan-Chen) to 20 (HCZ D3Q19: five fields, two sides, pack and unpack) plane copies; as
// separate cudaMemcpyAsync nodes each costs several microseconds of copy-engine set-up on the step's critical path, and the
// peer-memory ring wants them as stores anyway (the destination may be the neighbour's mailbox).
struct SegTable {
    void *dst[24];
    const void *src[24];
    unsigned long long bytes[24];
    int n;
};

__global__ void __launch_bounds__(256) copy_segments_kernel(SegTable T, RingSync rs)
{
    ring_kernel_begin(rs);
    const int sgm = blockIdx.y;
    if (sgm < T.n) {
        char *d = (char *)T.dst[sgm];
        const char *s = (const char *)T.src[sgm];
        const unsigned long long nb = T.bytes[sgm];
        const unsigned long long t0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (unsigned long long)gridDim.x * blockDim.x;
        if ((((unsigned long long)d | (unsigned long long)s | nb) & 15ull) == 0) {
            const unsigned long long n16 = nb >> 4;
            const int4 z = make_int4(0, 0, 0, 0);
            for (unsigned long long i = t0; i < n16; i += stride) ((int4 *)d)[i] = s ? ((const int4 *)s)[i] : z;
        } else if ((((unsigned long long)d | (unsigned long long)s | nb) & 7ull) == 0) {   // fp64 planes of an odd row count
            const unsigned long long n8 = nb >> 3;
            for (unsigned long long i = t0; i < n8; i += stride) ((long long *)d)[i] = s ? ((const long long *)s)[i] : 0ll;
        } else {
            for (unsigned long long i = t0; i < nb; i += stride) d[i] = s ? s[i] : (char)0;
        }
    }
    ring_kernel_end(rs);
}

struct SegList {
    SegTable T;
    SegList() { T.n = 0; }
    // a full table is flushed by the caller through launch(); returns false when there is no room
    bool add(void *dst, const void *src, size_t bytes)
    {
        if (T.n >= 24) return false;
        T.dst[T.n] = dst; T.src[T.n] = src; T.bytes[T.n] = bytes; ++T.n;
        return true;
    }
    // sync_mode 1: the pack of `phase` (signal when done), 2: its unpack (wait first); only honoured while the ring is fused
    int launch(clbm_ctx *c, const char *name, int phase = 0, int sync_mode = 0)
    {
        if (T.n == 0) return 0;
        unsigned long long mx = 0;
        for (int i = 0; i < T.n; ++i) mx = T.bytes[i] > mx ? T.bytes[i] : mx;
        long long blocks = (long long)((mx / 16 + 255) / 256);
        if (blocks < 1) blocks = 1;
        if (blocks > 148) blocks = 148;          // a grid-stride loop: one block per SM and segment is plenty for a few MB
        LaunchScope ls(c, name);
        const RingSync rs = ring_sync_for(c, phase, sync_mode, (unsigned)(blocks * T.n));
        copy_segments_kernel<<<dim3((unsigned)blocks, T.n), 256, 0, c->stream>>>(T, rs);
        CLBM_CUDA(cudaGetLastError());
        T.n = 0;
        return 0;
    }
};

int halo_pack(clbm_ctx *c, int phase)
{
    const Geom &g = c->geo;
    const size_t pl = (size_t)g.plane;
    c->launches += 0;
    if (phase == 0) {
        HaloField hf[8];
        const int nf = phase0_fields(c, hf);
        const bool hcz3 = c->prm.model == CLBM_MODEL_HCZ_D3Q19;
        if (c->halo0_packed) {   // the boundary-moment kernel of this stage stored into the send blocks itself
            const int how = c->halo0_packed;
            c->halo0_packed = 0;
            if (how == 2 || !c->ring_fuse) return 0;     // (a fused ring needs a kernel whose last block signals: that kernel, or the copy below)
        }
        SegList L;
        for (int side = 0; side < 2; ++side) {
            double *dst = (double *)halo_send_ptr(c, 0, side);
            for (int i = 0; i < nf; ++i) {
                const int d = hf[i].depth;
                const int x0 = side ? g.nx - d : 0;
                if (hcz3 && c->sweep_active && hf[i].fld == 0) {
                    // phi of the sweep kernel = node array + edge sums: folded while packing (the neighbour's ghost planes hold plain values)
                    if (int rc = hcz3d_pack_phi_merged(c, dst, x0, d)) return rc;
                } else {
                    const double *src = hcz3 ? hcz3d_moment_array(c, hf[i].fld) : c->fld[hf[i].fld];
                    if (!L.add(dst, src + (size_t)(x0 + g.G) * pl, d * pl * sizeof(double))) { set_error("halo segment table full"); return CLBM_ESTATE; }
                }
                dst += d * pl;
            }
        }
        return L.launch(c, "pack_moment_halo", 0, 1);
    }
    if (phase == 1) {
        // the ghost planes the push wrote across the slab face are contiguous per direction: 2 x ncross x sets plane copies in one
        // vectorised launch (one thread per double and a fence behind each: 50 us for a 512 x 512 Shan-Chen face)
        const CrossTable T = cross_table(c);
        SegList L;
        for (int side = 0; side < 2; ++side) {
            const int xg = side ? g.nx : -1;
            for (int sl = 0; sl < T.ncross * T.sets; ++sl) {
                const int s = sl / T.ncross, k = T.ks[side][sl % T.ncross];
                if (!L.add(T.send[side] + (size_t)sl * pl, T.pop[s] + (size_t)k * g.ncs + (size_t)(xg + g.G) * pl, pl * sizeof(double))) { set_error("halo segment table full"); return CLBM_ESTATE; }
            }
        }
        return L.launch(c, "pack_cross", 1, 1);
    }
    if (phase == 2) {
        SegList L;
        for (int side = 0; side < 2; ++side) {
            const int x0 = side ? g.nx - g.G : 0;
            L.add(halo_send_ptr(c, 2, side), c->flag + (size_t)(x0 + g.G) * pl, (size_t)g.G * pl);
        }
        return L.launch(c, "pack_mask_halo", 2, 1);
    }
    set_error("bad halo phase %d", phase);
    return CLBM_EINVAL;
}

// crossing populations land in the boundary plane only where the sender really wrote them:
// the upstream node (in the neighbour's slab = our ghost plane) is fluid and the target is not a wall;
// elsewhere the slot was filled locally by the target's own half-way bounce-back.
int halo_unpack(clbm_ctx *c, int phase)
{
    const Geom &g = c->geo;
    const size_t pl = (size_t)g.plane;
    if (phase == 0) {
        if (c->halo0_direct && c->ring_fuse != 1) return 0;   // the neighbours stored into our ghost planes themselves
        HaloField hf[8];
        const int nf = phase0_fields(c, hf);
        SegList L;
        for (int side = 0; side < 2; ++side) {
            const double *src = (const double *)c->halo[0][side][1];
            for (int i = 0; i < nf; ++i) {
                const int d = hf[i].depth;
                const int x0 = side ? g.nx : -d;   // ghost planes on that side
                double *dstf = c->prm.model == CLBM_MODEL_HCZ_D3Q19 ? hcz3d_moment_array(c, hf[i].fld) : c->fld[hf[i].fld];
                if (!L.add(dstf + (size_t)(x0 + g.G) * pl, src, d * pl * sizeof(double))) { set_error("halo segment table full"); return CLBM_ESTATE; }
                src += d * pl;
            }
        }
        return L.launch(c, "unpack_moment_halo", 0, 2);
    }
    if (phase == 1) {
        // data received from the side-0 neighbour moves in +x (c_x = +1) into plane 0, and vice versa
        LaunchScope ls(c, "unpack_cross");
        const CrossTable T = cross_table(c);
        long long bx = grid_for(g.plane, 256);
        if (bx > 148 * 4) bx = 148 * 4;
        dim3 grid((unsigned)bx, 2);
        const RingSync rs = ring_sync_for(c, 1, 2, grid.x * grid.y);
        if (c->Q == 9) unpack_cross_kernel<D2Q9><<<grid, 256, 0, c->stream>>>(T, c->flag, g, rs);
        else unpack_cross_kernel<D3Q19><<<grid, 256, 0, c->stream>>>(T, c->flag, g, rs);
        CLBM_CUDA(cudaGetLastError());
        return 0;
    }
    if (phase == 2) {
        SegList L;
        for (int side = 0; side < 2; ++side) {
            const int x0 = side ? g.nx : -g.G;
            L.add(c->flag + (size_t)(x0 + g.G) * pl, c->halo[2][side][1], (size_t)g.G * pl);
        }
        return L.launch(c, "unpack_mask_halo", 2, 2);
    }
    set_error("bad halo phase %d", phase);
    return CLBM_EINVAL;
}

// ---- clbm_upload: populations of bounce_back nodes in the buffer the caller did NOT select -------------------------
// Only bounce_back nodes keep what they were initialised with (every slot of a bulk node is rewritten by each step), so the
// reference's second buffer matters exactly there (PF/apps/twoLayeredFlow2D.h:184-187 fills both buffers at init).
__global__ void __launch_bounds__(256)
scatter_nodes_kernel(double *__restrict__ pop, const long long *__restrict__ idx, const double *__restrict__ vals, long long nn, int Q, long long ncs)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nn) return;
    const long long i = idx[j];
    for (int k = 0; k < Q; ++k) pop[(size_t)k * ncs + i] = vals[(size_t)k * nn + j];
}

// vals: [sets][Q][nn] host-ordered values of the nodes idx[] (storage cell indices) for device buffer `buffer`
int scatter_node_pops(clbm_ctx *c, int buffer, const long long *idx_dev, const double *vals_dev, long long nn)
{
    for (int s = 0; s < c->sets; ++s) {
        LaunchScope ls(c, "upload_scatter_nodes");
        scatter_nodes_kernel<<<grid_for(nn, 256), 256, 0, c->stream>>>(c->pop[s][buffer], idx_dev, vals_dev + (size_t)s * c->Q * nn, nn, c->Q, c->geo.ncs);
        CLBM_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace clbm
