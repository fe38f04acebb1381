// clbm_internal.h -- context object behind the C ABI (include/clbm.h).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/clbm.h"
#include "lattice.cuh"

namespace clbm {

// scalar model parameters handed to kernels by value
struct ModelParams {
    double omega, gravity;
    double rho_w, a, b, R, TT;                    // Shan-Chen / Yuan-CS
    double phi_l, phi_g, rho_l, rho_g, kappa;     // HCZ
    int sc_force;
    // derived on the host with the reference's expressions (clbm_create)
    double tau;                  // 1/omega
    double psiw_pos, psiw_neg;   // wall pseudopotential for G1 = +1/3 / -1/3 at the centre node
    double gx, gy, G, p_shift;   // Shan-Chen constant-G variant (SC/apps/twoLayeredFlow2D.h:150-153)
    double gx_const;             // HCZ layered variant: constant x force (PF/apps/twoLayeredFlow2D.h:128); sc_force carries the variant
    double kpsi;                 // 2 / (|G| cs2): psi = sqrt(kpsi * (rho/3 - P_eos - p_shift))
    double inv_dphi, drho;       // HCZ total_rho: rho = rho_g + (phi - phi_g) * inv_dphi * drho, inv_dphi = 1/(phi_l - phi_g)
    double s_e, s_eps, s_q;      // MRT rates of the non-hydrodynamic moments (CLBM_COLLISION_MRT, HCZ D2Q9; mrt.cuh)
};

// scalar parameters of the kernels from the C-ABI parameter block, derived with the reference's own expressions
inline void derive_model_params(const clbm_params *p, ModelParams &m)
{
    m.omega = p->omega; m.gravity = p->gravity;
    m.rho_w = p->rho_w; m.a = p->a; m.b = p->b; m.R = p->R; m.TT = p->TT;
    m.phi_l = p->phi_l; m.phi_g = p->phi_g; m.rho_l = p->rho_l; m.rho_g = p->rho_g; m.kappa = p->kappa;
    m.sc_force = p->sc_force;
    m.tau = 1. / p->omega;
    m.inv_dphi = (p->phi_l != p->phi_g) ? 1.0 / (p->phi_l - p->phi_g) : 0.0;
    m.drho = p->rho_l - p->rho_g;
    m.s_e = p->s_e; m.s_eps = p->s_eps; m.s_q = p->s_q;
    {
        // wall pseudopotential: laplace2D.h:210 evaluates psi_yuan_from_rho(rho_w) (own branch G1(rho_w));
        // contactAngle2D.h:259-262 re-evaluates it on the CENTRE node's branch G1c = +-1/3
        const double cs2 = 1.0 / 3.0, rw = p->rho_w, dw = (1.0 - rw);
        const double Zw = 1.0 + (4.0 * rw - 2.0 * rw * rw) / (dw * dw * dw);
        m.gx = p->gx; m.gy = p->gy; m.G = p->G; m.p_shift = p->p_shift; m.gx_const = p->gx_const;
        m.kpsi = (p->G != 0.0) ? 2.0 / (fabs(p->G) * cs2) : 0.0;
        if (p->sc_force == CLBM_SC_FORCE_CONSTG) {
            // psi_w = psi_from_rho(rho_w) with the same constant-G mapping (twoLayeredFlow2D.h:226)
            const double Pw = rw * p->R * p->TT * Zw - p->a * rw * rw + p->p_shift;
            const double Sw = cs2 * rw - Pw;
            m.psiw_pos = m.psiw_neg = (Sw <= 0.0) ? 0.0 : sqrt(2.0 * Sw / (fabs(p->G) * cs2));
        } else if (p->sc_force == CLBM_SC_FORCE_CONTACT) {
            const double vp = 6.0 * rw * (p->R * p->TT * Zw - p->a * rw - cs2) / cs2;
            const double vn = 6.0 * rw * (p->R * p->TT * Zw - p->a * rw - cs2) / -cs2;
            m.psiw_pos = (vp > 0.0) ? sqrt(vp) : 0.0;
            m.psiw_neg = (vn > 0.0) ? sqrt(vn) : 0.0;
        } else {
            const double Pw = rw * p->R * p->TT * Zw - p->a * rw * rw;
            const double sw = p->R * p->TT * Zw - p->a * rw - cs2;
            const double G1w = (sw > 0.0) ? cs2 : -cs2;
            const double vw = 6.0 * (Pw - cs2 * rw) / G1w;
            m.psiw_pos = m.psiw_neg = (vw > 0.0) ? sqrt(vw) : 0.0;
        }
    }
}

struct KernelTiming {
    std::string name;
    float ms;
};

// tuning knobs from the environment, read ONCE per context in clbm_create (-1 = not set): the launch paths never call getenv
struct EnvKnobs {
    int sc_xchunk, sc_tile, sc_cluster, tma_promo, sc_multi, sc2d_tma;
    int hcz_tile, hcz_xchunk, hcz2d_tile, hcz2d_xchunk, hcz2d_multi;
    int hcz3d_sweep;   // 1 / 0: force / forbid the single-sweep HCZ D3Q19 kernel (default: where eligible)
    int hcz3d_sweep_var;   // load-issue variant of that kernel (hcz3d_sweep.cu, bit-identical results)
    int hcz3d_sweep_ko;    // knock-out bits for timing experiments (WRONG results when set; tools/hcz3d_sweep_variants.py)
    int slab_graph;    // 0: never capture the slab step in a CUDA graph
    int persist;       // 0: never use the persistent multi-step kernels of the L2-resident lattices
    int ring_fuse;     // peer ring: 0 separate signal / wait kernels, 1 both fused into the pack / unpack kernels (measured slower), default (2): signals fused only
    int slab_overlap;  // 0: sequential slab protocol only; 1: interior-first overlap; 2: halo-first overlap (default: per model, clbm_api.cu)
    int force_slab;    // 1: treat a full-width lattice as an x-slab (a ring of ONE context, its own neighbour: tests of the ring code)
};
inline int env_int(const char *name, int unset = -1)
{
    const char *e = getenv(name);
    return (e && *e) ? atoi(e) : unset;
}
inline void read_env_knobs(EnvKnobs &k)
{
    k.sc_xchunk = env_int("CLBM_SC_XCHUNK");
    k.sc_tile = env_int("CLBM_SC_TILE");
    k.sc_cluster = env_int("CLBM_SC_CLUSTER");
    k.sc_multi = env_int("CLBM_SC_MULTI");
    k.sc2d_tma = env_int("CLBM_SC2D_TMA");   // 0: off, 1: force the default shape, 2..7: other tile / stage shapes
    k.tma_promo = env_int("CLBM_TMA_PROMO");
    k.hcz_tile = env_int("CLBM_HCZ_TILE");
    k.hcz_xchunk = env_int("CLBM_HCZ_XCHUNK");
    k.hcz2d_multi = env_int("CLBM_HCZ2D_MULTI");   // 1: clbm_step(n >= 2) of an HCZ D2Q9 lattice as one cooperative launch (opt-in: measured slower at configs[1])
    k.hcz2d_tile = env_int("CLBM_HCZ2D_TILE");
    k.hcz2d_xchunk = env_int("CLBM_HCZ2D_XCHUNK");
    k.hcz3d_sweep = env_int("CLBM_HCZ3D_SWEEP");
    k.hcz3d_sweep_var = env_int("CLBM_HCZ3D_SWEEP_VAR");
    k.hcz3d_sweep_ko = env_int("CLBM_HCZ3D_SWEEP_KO");
    k.slab_graph = env_int("CLBM_SLAB_GRAPH");
    k.persist = env_int("CLBM_PERSIST");
    k.ring_fuse = env_int("CLBM_RING_FUSE");
    k.slab_overlap = env_int("CLBM_SLAB_OVERLAP");
    k.force_slab = env_int("CLBM_FORCE_SLAB");
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE property of a kernel: one bit per device ordinal,
// so a second device used by the same process gets its own call (a per-process flag left it unset there)
struct PerDeviceOnce {
    std::atomic<unsigned long long> done{0ull};
    bool need(int dev) const { return dev < 0 || dev >= 64 || !((done.load(std::memory_order_acquire) >> dev) & 1ull); }
    void mark(int dev) { if (dev >= 0 && dev < 64) done.fetch_or(1ull << dev, std::memory_order_release); }
};

// an encoded tensor map, cached per (buffer, box shape, L2 promotion): encoding costs microseconds of host time per launch,
// which is what a slab step at strong-scaling sizes does not have
struct TmapEntry {
    const void *base;
    unsigned box[4];
    int promo;
    alignas(64) unsigned char map[128];   // CUtensorMap
};

}  // namespace clbm

struct clbm_ctx {
    clbm_params prm;
    clbm::Geom geo;
    clbm::ModelParams mp;
    int Q, sets, device;
    int parity;           // which device buffer is "in"
    int host_parity0;     // parity value the host uploaded
    long long steps_taken;
    int multi;            // 1: x-slab of a wider lattice (ghost planes filled by exchange)
    cudaStream_t stream;      // launching stream (all work of a single slab; the interior of an overlapped slab step)
    cudaStream_t stream_u;    // second stream of clbm_upload (two host-to-device copies in flight)
    cudaStream_t stream_b;    // boundary stream of the overlap protocol (high priority): boundary planes, pack/unpack, exchange
    cudaEvent_t ev_main, ev_b;   // interior done / boundary + exchange done (cross-stream ordering between steps)
    cudaEvent_t ev0, ev1;
    int64_t launches;

    // populations: pop[set][buffer], each Q * geo.ncs doubles
    double *pop[2][2];
    uint8_t *flag;        // geo.ncs
    // moment / stage fields, geo.ncs doubles each (meaning depends on the model)
    double *fld[12];
    int nfld;
    // single-sweep HCZ D3Q19 step (hcz3d_sweep.cu): two sets of moment arrays (mom[0] = fld[0..4]) and of edge arrays, the
    // set holding the moments of the current "in" populations, and whether it is valid (an upload, a device-side init, a
    // direct moments pass or a step of another kernel path invalidates it: the next step then rebuilds it from the populations)
    double *mom[2][5], *mome[2][5];
    int mom_src, mom_valid;
    int sweep_active;             // x-slab only: the current step runs the sweep kernel, the moment halo comes from mom[mom_src] (+ edges)
    int walls_known, has_walls;   // result of the one-time scan for bounce_back nodes (the sweep kernel is wall-free only)
    // reduction scratch
    double *red_dev;
    double *red_host;     // pinned
    // halo buffers: [phase][side][send=0/recv=1]
    void *halo[3][2][2];
    size_t halo_bytes[3];
    void *mailbox;              // the one allocation behind halo[][][] + a page of flag words (support_kernels.cu: halo_alloc)
    size_t mailbox_bytes, mailbox_flags_off;
    // library-driven ring (slab_comm.cu): ncclComm_t of this rank, its rank and the ring size (0 = no communicator)
    void *comm;
    int comm_rank, comm_size;
    // peer-memory ring (slab_comm.cu): 0 = none, 1 = neighbours' mailboxes mapped through CUDA IPC (one process per GPU),
    // 2 = neighbours are contexts of this process.  peer_base[side] = mailbox of the neighbour on that side.
    int peer_mode;
    void *peer_base[2];
    int *peer_err;              // pinned + mapped: a wait kernel that timed out writes its phase + 1 here
    // Shan-Chen x-slabs: the psi field lives INSIDE the mailbox allocation, so that on a peer ring a neighbour stores its boundary
    // psi plane straight into our ghost plane (no receive block, no unpack launch for the moment halo)
    int fld0_in_mailbox;
    size_t mailbox_psi_off;     // byte offset of fld[0] in the mailbox (0: not there)
    int halo0_direct;           // peer ring connected with direct ghost planes: halo_send_ptr(0, side) is the neighbour's ghost plane
    int peer_nx[2];             // local plane count of the neighbour on each side (its right ghost plane is plane nx + G of its storage)
    int halo0_packed;           // the moment kernels of this stage wrote the phase-0 send blocks themselves (Shan-Chen boundary psi)
    int ring_fuse;              // 1 while clbm_slab_step issues stages whose pack / unpack kernels carry the signal / wait themselves
    void *slab_graph[2];        // cudaGraphExec_t of two consecutive slab steps starting at parity 0 / 1
    int64_t slab_graph_launches[2];   // kernels one replay launches (counted while capturing)
    int slab_graph_failed;
    // staging for host<->device slab transfers (pinned), grown on demand
    void *stage;
    size_t stage_bytes;
    // dominant-kernel timing (event pairs around the collide/stream launches)
    bool ktiming;
    int ktiming_cap;
    std::vector<cudaEvent_t> kev;
    const char *kname;
    // profiling
    bool profiling;
    std::vector<clbm::KernelTiming> prof;
    clbm::EnvKnobs env;
    std::vector<clbm::TmapEntry> tmaps;
    // work queues of the persistent Shan-Chen D3Q19 kernel (sc_fused_tma_persist.cu): 8 x {next item, CTAs done}, one pair per
    // launch in rotation (two launches of the overlap protocol may run at the same time); each returns to zero by itself
    int *sc_queue;
    int sc_queue_next;
    int sm_count;
    // column-resident multi-step kernel of the L2-resident D2Q9 lattices (sc_fused.cu): steps completed per column, counting
    // across launches
    int *resident_progress;
    int resident_epoch;
    // persistent device scratch of clbm_download_fields / clbm_download_force (grown on demand, freed in clbm_destroy)
    double *scratch;
    size_t scratch_bytes;
};

namespace clbm {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define CLBM_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return clbm::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

// RAII-ish helper that counts a launch and (when profiling) brackets it with events
struct LaunchScope {
    clbm_ctx *c;
    const char *name;
    cudaEvent_t a, b;
    bool timed;
    LaunchScope(clbm_ctx *ctx, const char *n, bool dominant = false);
    ~LaunchScope();
};

// model steps (one full time step on ctx->stream, x-planes [0,nx)); defined per model
int sc_step(clbm_ctx *c);
int hcz2d_step(clbm_ctx *c);
int hcz3d_step(clbm_ctx *c);
// staged halves for the slab exchange protocol
int model_stage(clbm_ctx *c, int stage);
// macroscopic fields into device arrays of nx*ny*nz (no ghosts); NULL = skip
int model_fields(clbm_ctx *c, double *s0, double *s1, double *s2, double *ux, double *uy, double *uz);
int model_reduce(clbm_ctx *c, int kind, double *out);
int model_init_case(clbm_ctx *c, int case_id, const double *args, int nargs);
int halo_pack(clbm_ctx *c, int phase);
int halo_unpack(clbm_ctx *c, int phase);
int field_scratch(clbm_ctx *c, size_t bytes, double **out);   // persistent device scratch of the downloads / reductions
int halo_alloc(clbm_ctx *c);
// HCZ D3Q19: node array of moment m the halo exchange reads / fills (the sweep kernel's current set, else fld[m]), and the
// pack of the phi halo with the sweep kernel's edge sums folded in (hcz3d_sweep.cu)
double *hcz3d_moment_array(const clbm_ctx *c, int m);
int hcz3d_pack_phi_merged(clbm_ctx *c, double *dst, int x0, int nplanes);
size_t halo_block_offset(const clbm_ctx *c, int phase, int side, int recv);
void *halo_send_ptr(const clbm_ctx *c, int phase, int side);
int scatter_node_pops(clbm_ctx *c, int buffer, const long long *idx_dev, const double *vals_dev, long long nn);

inline int grid_for(long long n, int block) { return (int)((n + block - 1) / block); }

}  // namespace clbm
