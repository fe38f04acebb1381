// sc_kernels.cu -- Shan-Chen (Yuan-CS) time step, staged form:
//   sc_psi_kernel     : rho = sum_k f_k, psi(rho) per node            (level 0/1 of SURVEY.md A.7)
//   sc_collide_kernel : force from the psi stencil, BGK collision, push streaming with
//                       half-way bounce-back (SC/apps/laplace2D.h:285-306, :260-270)
// The fused plane-marching form lives in sc_fused.cu; both share sc_cell.cuh.
#include "ring_sync.cuh"
#include "sc_cell.cuh"

namespace clbm {

RingSync ring_sync_for(const clbm_ctx *c, int phase, int mode, unsigned nblocks);   // slab_comm.cu

template <class L, bool GUO = false>
__global__ void __launch_bounds__(256)
sc_psi_kernel(const double *__restrict__ fin, const uint8_t *__restrict__ flag, double *__restrict__ psi,
              Geom g, ModelParams mp, int x0, long long ncell)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell) return;
    const long long i = (long long)(x0 + g.G) * g.plane + t;
    double f[L::Q];
#pragma unroll
    for (int k = 0; k < L::Q; ++k) f[k] = fin[(size_t)k * g.ncs + i];
    // |value| = psi(rho), sign bit = branch of G1 (set: G1 = -1/3); walls hold +0
    double v = 0.0;
    if (flag[i] != CELL_BB) {
        if constexpr (GUO) {
            // Rayleigh-Taylor variant: psi = 1 - exp(-rho) >= 0, no G1 branch; wall nodes carry zero populations in the
            // reference (inigeom zeroes them, nothing writes them), i.e. psi(wall) = 0, which is what +0 stands for here
            v = scrt_psi(Mom<L>::sum(f));
        } else {
            bool g1_pos;
            const double ps = sc_psi_g1(mp, Mom<L>::sum(f), g1_pos);
            v = g1_pos ? ps : -ps;
        }
    }
    psi[i] = v;
}

// psi of the two boundary planes of a slab in ONE launch (blockIdx.y = side), written to the psi field AND straight into the
// send block of that side -- on a peer ring the neighbour's mailbox -- so that the moment halo of a Shan-Chen slab step needs
// no pack launch of its own (two of the ten small dependent kernels of a sequential step)
template <class L, bool GUO = false>
__global__ void __launch_bounds__(256)
sc_psi_boundary_kernel(const double *__restrict__ fin, const uint8_t *__restrict__ flag, double *__restrict__ psi, Geom g, ModelParams mp,
                       double *__restrict__ send0, double *__restrict__ send1, RingSync rs)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < g.plane) {
        const int side = blockIdx.y;
        const long long i = (long long)((side ? g.nx - 1 : 0) + g.G) * g.plane + t;
        double f[L::Q];
#pragma unroll
        for (int k = 0; k < L::Q; ++k) f[k] = fin[(size_t)k * g.ncs + i];
        double v = 0.0;
        if (flag[i] != CELL_BB) {
            if constexpr (GUO) {
                v = scrt_psi(Mom<L>::sum(f));
            } else {
                bool g1_pos;
                const double ps = sc_psi_g1(mp, Mom<L>::sum(f), g1_pos);
                v = g1_pos ? ps : -ps;
            }
        }
        psi[i] = v;
        (side ? send1 : send0)[t] = v;
    }
    ring_kernel_end(rs);   // fused ring: the last block tells both neighbours that their ghost psi planes are complete
}

template <class L, bool GUO = false, bool MRT = false>
__global__ void __launch_bounds__(256, 2)
sc_collide_kernel(const double *__restrict__ fin, double *__restrict__ fout, const uint8_t *__restrict__ flag,
                  const double *__restrict__ psi, Geom g, ModelParams mp, int x0, long long ncell)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell) return;
    const int x = x0 + (int)(t / g.plane);
    const int r = (int)(t % g.plane);
    const int y = r / g.nz, z = r % g.nz;
    const Nbr n = make_nbr(g, x, y, z);
    if (flag[n.i] != CELL_BULK) return;

    double f[L::Q], out[L::Q];
#pragma unroll
    for (int k = 0; k < L::Q; ++k) f[k] = fin[(size_t)k * g.ncs + n.i];

    ScForceSums s = {{0., 0., 0.}, {0., 0., 0.}, 0u};
    sc_gather_force<L, GUO>(s, n, flag, psi);
    const double pc = psi[n.i];
    if constexpr (GUO && MRT) scrt_collide_mrt<L>(mp, f, s, Mom<L>::sum(f), pc, out);
    else if constexpr (GUO) scrt_collide<L>(mp, f, s, Mom<L>::sum(f), pc, out);
    else if constexpr (MRT) sc_collide_mrt<L>(mp, f, s, Mom<L>::sum(f), fabs(pc), !signbit(pc), out);
    else sc_collide<L>(mp, f, s, fabs(pc), !signbit(pc), out);

#pragma unroll
    for (int k = 0; k < L::Q; ++k) {
        if (k == L::REST) { fout[(size_t)k * g.ncs + n.i] = out[k]; continue; }
        if (s.wall & (1u << k)) fout[(size_t)L::opp(k) * g.ncs + n.i] = out[k];   // half-way bounce-back
        else fout[(size_t)k * g.ncs + n.at<L>(k)] = out[k];
    }
}

template <class L, bool GUO = false>
__global__ void __launch_bounds__(256)
sc_fields_kernel(const double *__restrict__ fin, const uint8_t *__restrict__ flag, const double *__restrict__ psi,
                 Geom g, ModelParams mp, double *s0, double *s1, double *ux, double *uy, double *uz, double *fx, double *fy,
                 double *fz, long long ncell)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell) return;
    const int x = (int)(t / g.plane);
    const int r = (int)(t % g.plane);
    const int y = r / g.nz, z = r % g.nz;
    const Nbr n = make_nbr(g, x, y, z);
    double f[L::Q];
#pragma unroll
    for (int k = 0; k < L::Q; ++k) f[k] = fin[(size_t)k * g.ncs + n.i];
    double rho = Mom<L>::sum(f), pr = 0.0, u[3] = {0., 0., 0.}, F[3] = {0., 0., 0.};
    if (flag[n.i] == CELL_BULK) {
        ScForceSums s = {{0., 0., 0.}, {0., 0., 0.}, 0u};
        sc_gather_force<L, GUO>(s, n, flag, psi);
        if constexpr (GUO) scrt_outputs<L>(mp, f, s, rho, pr, u, F);
        else sc_outputs<L>(mp, f, s, rho, pr, u, F);
    }
    if (s0) s0[t] = rho;
    if (s1) s1[t] = pr;
    if (ux) ux[t] = u[0];
    if (uy) uy[t] = u[1];
    if (uz) uz[t] = u[2];
    if (fx) fx[t] = F[0];
    if (fy) fy[t] = F[1];
    if (fz) fz[t] = F[2];
}

// ---- host side ---------------------------------------------------------------------------
// Rayleigh-Taylor variant (psi = 1 - exp(-rho), Guo forcing): D2Q9 only, enforced by clbm_create
static bool is_guo(const clbm_ctx *c) { return c->mp.sc_force == CLBM_SC_FORCE_EXPGUO; }
template <class L> static int sc_psi_range(clbm_ctx *c, int x0, int x1)
{
    const long long n = (long long)(x1 - x0) * c->geo.plane;
    if (n <= 0) return 0;
    LaunchScope ls(c, "sc_psi");
    if (L::D == 2 && is_guo(c)) sc_psi_kernel<D2Q9, true><<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->flag, c->fld[0], c->geo, c->mp, x0, n);
    else sc_psi_kernel<L><<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->flag, c->fld[0], c->geo, c->mp, x0, n);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}
template <class L> static int sc_collide_range(clbm_ctx *c, int x0, int x1)
{
    const long long n = (long long)(x1 - x0) * c->geo.plane;
    if (n <= 0) return 0;
    LaunchScope ls(c, "sc_collide_stream", true);
    if (L::D == 2 && c->prm.collision == CLBM_COLLISION_MRT && is_guo(c))
        sc_collide_kernel<D2Q9, true, true><<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->pop[0][1 - c->parity], c->flag,
                                                                                    c->fld[0], c->geo, c->mp, x0, n);
    else if (c->prm.collision == CLBM_COLLISION_MRT)
        sc_collide_kernel<L, false, true><<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->pop[0][1 - c->parity], c->flag,
                                                                                     c->fld[0], c->geo, c->mp, x0, n);
    else if (L::D == 2 && is_guo(c))
        sc_collide_kernel<D2Q9, true><<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->pop[0][1 - c->parity], c->flag,
                                                                              c->fld[0], c->geo, c->mp, x0, n);
    else
        sc_collide_kernel<L><<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->pop[0][1 - c->parity], c->flag,
                                                                     c->fld[0], c->geo, c->mp, x0, n);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

int sc_fused_step(clbm_ctx *c);    // sc_fused.cu
int sc_fused_launch(clbm_ctx *c);
int sc_collide_all(clbm_ctx *c);

// psi of the two boundary planes only (what the neighbours need as moment halo when the fused kernel runs)
int sc_psi_boundary(clbm_ctx *c)
{
    const Geom &g = c->geo;
    double *s0 = (double *)halo_send_ptr(c, 0, 0), *s1 = (double *)halo_send_ptr(c, 0, 1);
    const dim3 grid(grid_for(g.plane, 256), 2);
    LaunchScope ls(c, "sc_psi_boundary");
    const RingSync rs = ring_sync_for(c, 0, 1, grid.x * grid.y);
    if (c->Q == 9 && is_guo(c)) sc_psi_boundary_kernel<D2Q9, true><<<grid, 256, 0, c->stream>>>(c->pop[0][c->parity], c->flag, c->fld[0], g, c->mp, s0, s1, rs);
    else if (c->Q == 9) sc_psi_boundary_kernel<D2Q9><<<grid, 256, 0, c->stream>>>(c->pop[0][c->parity], c->flag, c->fld[0], g, c->mp, s0, s1, rs);
    else sc_psi_boundary_kernel<D3Q19><<<grid, 256, 0, c->stream>>>(c->pop[0][c->parity], c->flag, c->fld[0], g, c->mp, s0, s1, rs);
    CLBM_CUDA(cudaGetLastError());
    c->halo0_packed = rs.mode ? 2 : 1;   // the halo_pack(0) that follows has nothing left to copy (2: nor to signal)
    return 0;
}
// collide + stream of the local planes in slab mode (ghost psi planes already unpacked); no parity flip
int sc_collide_slab(clbm_ctx *c)
{
    if (c->prm.fused) return sc_fused_launch(c);
    return sc_collide_all(c);
}

int sc_collide_all(clbm_ctx *c);
int sc_psi_all(clbm_ctx *c)
{
    return c->Q == 9 ? sc_psi_range<D2Q9>(c, 0, c->geo.nx) : sc_psi_range<D3Q19>(c, 0, c->geo.nx);
}
int sc_collide_all(clbm_ctx *c)
{
    return c->Q == 9 ? sc_collide_range<D2Q9>(c, 0, c->geo.nx) : sc_collide_range<D3Q19>(c, 0, c->geo.nx);
}

int sc_step(clbm_ctx *c)
{
    if (c->prm.fused && !c->multi) return sc_fused_step(c);
    int rc = sc_psi_all(c);
    if (rc) return rc;
    rc = sc_collide_all(c);
    if (rc) return rc;
    c->parity = 1 - c->parity;
    return 0;
}

static int sc_fields_all(clbm_ctx *c, double *s0, double *s1, double *ux, double *uy, double *uz, double *fx, double *fy, double *fz)
{
    // psi of the current populations (ghost planes of psi must already be valid in slab mode)
    int rc = sc_psi_all(c);
    if (rc) return rc;
    const long long n = (long long)c->geo.nx * c->geo.plane;
    LaunchScope ls(c, "sc_fields");
    if (c->Q == 9 && is_guo(c))
        sc_fields_kernel<D2Q9, true><<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->flag, c->fld[0], c->geo, c->mp, s0, s1, ux, uy, uz, fx, fy, fz, n);
    else if (c->Q == 9)
        sc_fields_kernel<D2Q9><<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->flag, c->fld[0], c->geo, c->mp, s0, s1, ux, uy, uz, fx, fy, fz, n);
    else
        sc_fields_kernel<D3Q19><<<grid_for(n, 256), 256, 0, c->stream>>>(c->pop[0][c->parity], c->flag, c->fld[0], c->geo, c->mp, s0, s1, ux, uy, uz, fx, fy, fz, n);
    CLBM_CUDA(cudaGetLastError());
    return 0;
}

int sc_fields(clbm_ctx *c, double *s0, double *s1, double *ux, double *uy, double *uz)
{
    return sc_fields_all(c, s0, s1, ux, uy, uz, nullptr, nullptr, nullptr);
}
// `force` of every node (SC/apps/laplace2D.h:198-242, contactAngle2D.h:248-293), 0 at non-bulk nodes
int sc_force_field(clbm_ctx *c, double *fx, double *fy, double *fz)
{
    return sc_fields_all(c, nullptr, nullptr, nullptr, nullptr, nullptr, fx, fy, fz);
}

}  // namespace clbm
