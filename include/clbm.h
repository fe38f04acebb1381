/*
 * clbm.h -- C ABI of the B200-native multiphase lattice-Boltzmann time step.
 *
 * This is the drop-in boundary for the one hot line of every CooLBM case file
 * in AmooMaD/Multiphase-LBM:
 *
 *     std::for_each(std::execution::par_unseq, lattice, lattice + dim.nelem, lbm);
 *     *parity = 1 - *parity;
 *
 * (reference: "shan-chen single component model/apps/laplace2D.h":506-507,
 *  "shan-chen single component model/apps/contactAngle2D.h":801-802,
 *  "Phase field model/apps/rayleighTaylor2D.h":980-983,
 *  "Phase field model/apps/laplace3D.h":943-946,
 *  "Abbashub LBM/apps/PulsatileBloodFlow2D.h":764-789).
 *
 * The reference has no FFI of its own: the functor `lbm` is an aggregate of raw
 * pointers (lattice, flag, parity) plus scalar model parameters passed by value.
 * The entry points below carry exactly that aggregate across a C boundary:
 * plain pointers, sizes and doubles; no C++ or torch types.
 *
 * Host arrays use the reference memory layout:
 *   lattice[s*2*npop + p*npop + k*nelem + i],  s = population set (0: f, 1: g),
 *   p = buffer selected by parity, k = direction, i = cell index,
 *   i = y + ny*x (2-D, y fastest)  /  i = z + nz*(y + ny*x) (3-D, z fastest),
 *   npop = Q*nelem;  flag[i] is uint8 {0: bounce_back, 1: bulk}.
 *
 * Every function returns 0 on success and a negative CLBM_E* code on failure;
 * clbm_last_error() returns the message of the last failure on this thread.
 * Nothing throws across the boundary.  There is no CPU fallback: without a
 * CUDA device clbm_create fails with CLBM_ENODEVICE.
 */
#ifndef CLBM_H
#define CLBM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLBM_ABI_VERSION 3 /* 3: + CLBM_SC_FORCE_EXPGUO, CLBM_CASE_SC_RT2D, clbm_diag_*, clbm_comm_* / clbm_slab_step; clbm_params grew at
                              its END (s_e, s_eps, s_q, collision, reserved0): a version-2 block zero-extended is a valid BGK block */

/* ---- models (one per reference functor family) -------------------------- */
#define CLBM_MODEL_SC_D2Q9    0 /* LBM_Laplace2D / LBM_contactAngle2D (Yuan-CS Shan-Chen) */
#define CLBM_MODEL_SC_D3Q19   1 /* same physics on the D3Q19 set of PF/apps/laplace3D.h:31-55 */
#define CLBM_MODEL_HCZ_D2Q9   2 /* LBM_rayleighTaylor2D (He-Chen-Zhang phase field) */
#define CLBM_MODEL_HCZ_D3Q19  3 /* LBM_laplace3D */
#define CLBM_MODEL_PULSATILE  4 /* LBM_PulsatileBloodFlow2D (pressure-based D2Q9 MRT) */

/* Shan-Chen force variant */
#define CLBM_SC_FORCE_LAPLACE 0 /* SC/apps/laplace2D.h:198-242: psi_w = psi(rho_w) on G1(rho_w); + gravity*rho in y */
#define CLBM_SC_FORCE_CONTACT 1 /* SC/apps/contactAngle2D.h:248-293: psi_w on the centre node's G1 branch; F=0 if rho<=0; no gravity */
#define CLBM_SC_FORCE_CONSTG  2 /* SC/apps/twoLayeredFlow2D.h:183-261: constant coupling G, psi = sqrt(2 (rho/3 - P_eos - p_shift) / (|G|/3)),
                                   psi_w = psi(rho_w), F=0 if rho<=0, uniform body force (gx, gy) added to F; pressure_node = P_eos */
#define CLBM_SC_FORCE_EXPGUO  3 /* SC/apps/RayleighTaylor2D.h (D2Q9 only): psi = 1 - exp(-rho) (:194-196), constant coupling G (the header's `g`),
                                   a bounce_back neighbour contributes the psi of the OPPOSITE neighbour (:246-262), + gravity*rho in y (:286);
                                   velocity shift u + F/(2 rho) (:343-351) and Guo's forcing term in the collision (:370-436); no wall force */

/* collision operator, `collision` member */
#define CLBM_COLLISION_BGK 0
#define CLBM_COLLISION_MRT 1 /* HCZ D2Q9 and Yuan-CS Shan-Chen D2Q9 (not CLBM_SC_FORCE_EXPGUO); BASELINE.json configs[1-2] ask for MRT, the reference functor is BGK:
                                parity of this operator is UNPINNED against the reference, pinned to BGK at S = omega I */

/* HCZ D2Q9 force variant, carried in the same `sc_force` member */
#define CLBM_HCZ_FORCE_GRAVITY 0 /* PF/apps/rayleighTaylor2D.h:316-337: F = kappa rho grad lap phi, + gravity*rho in y */
#define CLBM_HCZ_FORCE_LAYERED 1 /* PF/apps/twoLayeredFlow2D.h:310-330: F_x = kappa rho (grad lap phi)_x + rho*gx + gx_const, no y drive;
                                    the rest population uses grad lap RHO instead (:595-598) */

/* error codes */
#define CLBM_OK          0
#define CLBM_EINVAL     -1
#define CLBM_ENODEVICE  -2
#define CLBM_ECUDA      -3
#define CLBM_ENOMEM     -4
#define CLBM_ESTATE     -5

/* reductions, clbm_reduce(kind) */
#define CLBM_REDUCE_MASS    0 /* totalMass_*: sum of rho (SC) or phi (HCZ) over non-solid nodes (SC/apps/laplace2D.h:382-393) */
#define CLBM_REDUCE_ENERGY  1 /* computeEnergy_*: 0.5*sum_bulk(u.u)/nelem_global (SC/apps/laplace2D.h:368-380, PF/apps/rayleighTaylor2D.h:784-797) */
#define CLBM_REDUCE_UMAX    2 /* max |u| over bulk nodes (spurious-current diagnostic) */

/* built-in initial conditions, clbm_init_case(case_id) -- device-side equivalents of iniLattice + inigeom */
#define CLBM_CASE_SC_LAPLACE2D     0 /* SC/apps/laplace2D.h:132-145,397-404   args: {rhol, rhog, Rdrop}              */
#define CLBM_CASE_SC_CONTACT2D     1 /* SC/apps/contactAngle2D.h:126-137,442-455 args: {rhol, rhog, RR}              */
#define CLBM_CASE_SC_DROPLET3D     2 /* composed C4 case: sessile droplet, walls y=0,ny-1  args: {rhol, rhog, RR, yc} */
#define CLBM_CASE_SC_DROPLET3D_PER 3 /* fully periodic free droplet            args: {rhol, rhog, RR}                 */
#define CLBM_CASE_HCZ_RT2D         4 /* PF/apps/rayleighTaylor2D.h:155-193,802-820   args: none                       */
#define CLBM_CASE_HCZ_LAPLACE3D    5 /* PF/apps/laplace3D.h:170-213,830-849          args: none                       */
#define CLBM_CASE_SC_LAYERED2D     6 /* SC/apps/twoLayeredFlow2D.h:325-346,441-454 args: {rhol, rhog, h_lower, w_int}  */
#define CLBM_CASE_HCZ_LAYERED2D    7 /* PF/apps/twoLayeredFlow2D.h:148-196,737-757 args: {h_lower, w_int}              */
#define CLBM_CASE_SC_RT2D          8 /* SC/apps/RayleighTaylor2D.h:134-158,526-541  args: {rhol, rhog}                  */

typedef struct clbm_ctx clbm_ctx;

/*
 * Scalar members of the reference LBM_* aggregates.  Members a model does not
 * use are ignored.  The lattice handed to one context is an x-slab
 * [x_offset, x_offset+nx) of a global lattice nx_global wide (x is the slowest
 * index, so a slab is a contiguous range of every population array); a
 * single-GPU run has x_offset=0, nx_global=nx.
 */
typedef struct clbm_params {
    int32_t abi_version;    /* CLBM_ABI_VERSION */
    int32_t model;          /* CLBM_MODEL_* */
    int32_t nx, ny, nz;     /* local slab extent; nz = 1 for D2Q9 */
    int32_t nx_global;      /* global x extent (== nx on one GPU) */
    int32_t x_offset;       /* global x of local column 0 */
    int32_t sc_force;       /* CLBM_SC_FORCE_* */
    int32_t device;         /* CUDA device ordinal, -1 = current */
    int32_t fused;          /* 1: fused plane-marching kernels where available (default), 0: staged kernels */
    /* common */
    double omega;           /* BGK relaxation rate (all multiphase models) */
    double gravity;         /* body force in +y */
    /* Shan-Chen / Yuan-CS  (LBM_Laplace2D members, SC/apps/laplace2D.h:104-113) */
    double rho_w, a, b, R, TT;
    /* HCZ (LBM_rayleighTaylor2D members, PF/apps/rayleighTaylor2D.h:113-122) */
    double phi_l, phi_g, rho_l, rho_g, kappa;
    /* Shan-Chen constant-G variant (LBM members of SC/apps/twoLayeredFlow2D.h:150-153) */
    double gx, gy, G, p_shift;
    /* HCZ layered variant (LBM_twoLayeredPF2D members gx, Gx_const, PF/apps/twoLayeredFlow2D.h:127-128); gx is shared */
    double gx_const;
    /* collision operator (appended in ABI version 3; zero-initialised = BGK, the only operator the reference's SC / HCZ
     * functors have).  CLBM_COLLISION_MRT, D2Q9 models: relaxation in the moment basis of CooLBM_MRT_combustion.cpp:313-323
     * (rho, e, eps, jx, qx, jy, qy, pxx, pxy), rates S = (omega, s_e, s_eps, omega, s_q, omega, s_q, omega, omega) for BOTH
     * population sets, HCZ / Guo forcing term relaxed with (I - S/2) (the form of :2441, :2466); Yuan-CS Shan-Chen keeps its
     * tau-shifted equilibrium velocity (tau = 1/omega) and relaxes f - f_eq.  D3Q19 models (the reference has no D3Q19 basis):
     * the orthogonal basis of d'Humieres et al. 2002 (rho, e, eps, j_a, q_a, 3p_xx, 3pi_xx, p_ww, pi_ww, p_xy, p_yz, p_xz, m_a),
     * conserved and stress moments at omega, e at s_e, eps and pi at s_eps, q and m at s_q.  s_e = s_eps = s_q = omega is BGK. */
    double s_e, s_eps, s_q;
    int32_t collision;      /* CLBM_COLLISION_* */
    int32_t reserved0;
} clbm_params;

/* ---- life cycle ---------------------------------------------------------- */
int  clbm_create(const clbm_params *params, clbm_ctx **out);
int  clbm_destroy(clbm_ctx *ctx);
const char *clbm_last_error(void);
int  clbm_abi_version(void);

/* ---- state transfer (reference layout, host memory) ---------------------- */
/* replaces the host-side ownership of lattice/flag/parity: copies the parity-selected "in" buffer of every population
 * set, the bounce_back-node values of the other buffer and the mask to the device. */
int  clbm_upload(clbm_ctx *ctx, const double *lattice, const uint8_t *flag, int parity);
/* the same with a say on the buffer the parity does NOT select.  Every slot of a bulk node is rewritten by each step, so that
 * buffer matters only at bounce_back nodes, which keep their initial values for ever: clbm_upload (other_buffer = 1) hands
 * those node values over too -- the reference's layered HCZ init fills both buffers ("Phase field model/apps/
 * twoLayeredFlow2D.h":184-187) -- so that clbm_download_lattice later returns the reference's host array at EVERY node.
 * other_buffer = 0: lattice[] holds valid data only in the selected buffer (the other half may not even be mapped). */
int  clbm_upload2(clbm_ctx *ctx, const double *lattice, const uint8_t *flag, int parity, int other_buffer);
/* writes the current populations back into lattice[] (buffer selected by the returned
 * parity; the other buffer is left untouched) so unchanged host accessors keep working. */
int  clbm_download_lattice(clbm_ctx *ctx, double *lattice, int *parity);
/* macroscopic fields with the reference definitions; any pointer may be NULL.
 *   SC : s0 = density, s1 = pressure_node, u = u_actual
 *   HCZ: s0 = phi, s1 = total_P, s2 = total_rho, u = velocity
 * non-bulk nodes: s0/s2 as computed from the (zero) populations, s1 = 0, u = 0. */
int  clbm_download_fields(clbm_ctx *ctx, double *s0, double *s1, double *s2,
                          double *ux, double *uy, double *uz, uint8_t *flag);
/* Shan-Chen only: the interaction force `force(f0)` the reference's VTK writers print
 * (SC/apps/laplace2D.h:198-242, :356-364; contactAngle2D.h:248-293, :399-408); 0 at non-bulk nodes; NULL = skip */
int  clbm_download_force(clbm_ctx *ctx, double *fx, double *fy, double *fz);
/* device-side initial condition; args has the case-specific doubles listed above */
int  clbm_init_case(clbm_ctx *ctx, int case_id, const double *args, int nargs);

/* ---- the hot path --------------------------------------------------------- */
/* nsteps times { for_each(par_unseq, lattice, lattice+nelem, lbm); parity = 1-parity; }.
 * Asynchronous on the context's stream; any download/reduce synchronises. */
int  clbm_step(clbm_ctx *ctx, int nsteps);
int  clbm_sync(clbm_ctx *ctx);
/* same, bracketed by CUDA events on the launching stream; *ms = device time of the nsteps */
int  clbm_step_timed(clbm_ctx *ctx, int nsteps, float *ms);
/* number of kernel launches issued by this context so far */
int64_t clbm_launch_count(const clbm_ctx *ctx);
/* per-kernel device time of one step (CUDA events around every launch).
 * names/ms hold up to cap entries; returns the number of kernels or <0. */
int  clbm_profile_step(clbm_ctx *ctx, const char **names, float *ms, int cap);

/* average device time of the dominant kernel (the collide/stream sweep) over the launches issued since the
 * last call to clbm_kernel_timing_begin: CUDA event pairs recorded around each launch on the launching
 * stream.  At most cap_launches launches are sampled (the first ones).  *avg_ms <- mean, *count <- samples. */
int  clbm_kernel_timing_begin(clbm_ctx *ctx, int cap_launches);
int  clbm_kernel_timing_end(clbm_ctx *ctx, float *avg_ms, int *count, const char **kernel_name);

/* pinned host memory for the reference-layout arrays (so uploads/downloads run at full PCIe speed) */
int  clbm_alloc_host(size_t bytes, void **ptr);
int  clbm_free_host(void *ptr);

/* ---- diagnostics ---------------------------------------------------------- */
int  clbm_reduce(clbm_ctx *ctx, int kind, double *out);
/* calculateContactAngle's scans (SC/apps/contactAngle2D.h:465-529) on the device, Shan-Chen D2Q9, single slab:
 *   base_y = first non-solid row at x = 0 from y = 2 up (:473-476; >= ny-1: "no fluid row found", base = height = 0),
 *   base   = length of the run of rho > rho_cut on row base_y around x = nx/2 (:489-497),
 *   height = length of the run of fluid nodes with rho > rho_cut on column nx/2 from base_y up (:499-505).
 * The caller forms theta = atan((b/2)/(R-h)), R = (4h^2+b^2)/(8h) (:513-517).  NULL outputs are skipped. */
int  clbm_diag_contact_angle(clbm_ctx *ctx, double rho_cut, int *base_y, int *base, int *height);
/* the same scans on an x-slab, in global x, combined by the caller (two calls on every rank of the ring):
 *   1. base_y_in = -1:  out4[0] = base_y where this slab owns x = 0, ny elsewhere            -> MIN over the ranks
 *   2. base_y_in = that minimum: out4[1] = largest x < nx_global/2 on row base_y with rho <= rho_cut here (-1: none)  -> MAX
 *                                out4[2] = smallest x > nx_global/2 with rho <= rho_cut here (nx_global: none)          -> MIN
 *                                out4[3] = first y >= base_y on column nx_global/2 that ends the height run (ny where
 *                                          the slab does not own that column)                                          -> MIN
 *   base = max(0, out4[2] - out4[1] - 1), height = out4[3] - base_y; base_y >= ny-1: no fluid row, base = height = 0.
 * (clbm_diag_interface_heights needs no second form: a slab answers 0 for a column it does not own -> MAX over the ranks.) */
int  clbm_diag_contact_angle_slab(clbm_ctx *ctx, double rho_cut, int base_y_in, int *out4);
/* findInterfaceHeights' scans (PF/apps/rayleighTaylor2D.h:668-708) on the device, HCZ D2Q9 (on an x-slab: 0 for a column the
 * slab does not own, combine the ranks with MAX): the largest y in
 * [1, ny-2] with phi <= phi_mid on column x = 0 (the reference stores it in `bubble_y`) and on column x = nx/2 (`spike_y`);
 * 0 when the column has none (the reference initialises both ints from +-0.05). */
int  clbm_diag_interface_heights(clbm_ctx *ctx, double phi_mid, int *y_at_x0, int *y_at_xmid);

/* ---- x-slab ghost exchange (multi-GPU) ----------------------- */
/* A slab step is  clbm_step_begin (boundary planes, packs the send buffers) ->
 * caller moves send buffers to the neighbours' recv buffers (NCCL / P2P) ->
 * clbm_step_end (unpack + interior).  Buffers are device memory owned by ctx.
 * side: 0 = towards x-1 neighbour, 1 = towards x+1 neighbour.
 * phase: 0 = moment halo (rho / phi ...), 1 = crossing populations. */
int  clbm_halo_buffer(clbm_ctx *ctx, int phase, int side, int recv, void **dev_ptr, size_t *bytes);
int  clbm_halo_pack(clbm_ctx *ctx, int phase);
int  clbm_halo_unpack(clbm_ctx *ctx, int phase);
/* one time step split around the two exchanges (stage 20 = stage 0 with every moment rebuilt from the populations: the
 * call a driver makes before clbm_download_fields on a slab):
 *   stage 0: moments of the local slab + pack phase 0
 *   stage 1: unpack phase 0 + collide/stream + pack phase 1
 *   stage 2: unpack phase 1, flip parity                                         */
int  clbm_step_stage(clbm_ctx *ctx, int stage);
/* raw stream handle (cudaStream_t) so the caller can order its copies after ours */
void *clbm_stream(clbm_ctx *ctx);
/* Overlap protocol, available when clbm_overlap_supported() returns 1.  Two forms, clbm_overlap_variant():
 *  1 "interior first" ("boundary planes first on a high-priority stream -> start exchange -> interior on the main
 *    stream -> join"):
 *      stage 10: boundary stream: moments of the boundary planes + pack phase 0;  launching stream: collide/stream of the
 *                interior planes, which need nothing from the neighbours
 *      (caller exchanges phase 0 ON THE BOUNDARY STREAM)
 *      stage 11: boundary stream: unpack phase 0, collide/stream of the boundary planes, pack phase 1
 *      (caller exchanges phase 1 on the boundary stream)
 *      stage 12: boundary stream: unpack phase 1; the launching stream then waits for the boundary stream
 *  2 "halo first":
 *      stage 10: launching stream: moments of the boundary planes + pack phase 0
 *      (caller exchanges phase 0 ON THE LAUNCHING STREAM)
 *      stage 11: launching stream: unpack phase 0, collide/stream of a chunk of boundary planes per side, then of the interior;
 *                boundary stream (after the boundary chunks): pack phase 1
 *      (caller exchanges phase 1 on the boundary stream: it overlaps with the interior launch)
 *      stage 12: as above
 * CLBM_SLAB_OVERLAP = 0 / 1 / 2 (environment, read in clbm_create) turns the protocol off / forces a form.
 * clbm_boundary_stream returns the boundary stream (cudaStream_t), NULL when the protocol is not available. */
int  clbm_overlap_supported(const clbm_ctx *ctx);
int  clbm_overlap_variant(const clbm_ctx *ctx);
/* planes per side collided by the boundary launch (0: no separate interior launch): the interior launch, which the kernel
 * timing samples, covers nx - 2 * width planes */
int  clbm_overlap_width(const clbm_ctx *ctx);
void *clbm_boundary_stream(clbm_ctx *ctx);

/* ---- the ring driven from the library (one process per GPU, NCCL send/recv on the library's own streams) ----
 * The caller only transports a 128-byte ncclUniqueId once: rank 0 calls clbm_comm_unique_id, broadcasts the bytes by any
 * means (MPI, torch.distributed, a file), every rank calls clbm_comm_init on its slab context.  clbm_slab_step(ctx, n)
 * then runs n slab steps -- stages, both ghost exchanges, the overlap protocol where clbm_overlap_supported() -- without
 * returning to the caller and without a host synchronisation; every rank must call it with the same n.  NCCL is
 * resolved at run time (libnccl.so.2), libclbm.so does not link it.  nranks >= 2 (a single slab needs no ring). */
int  clbm_comm_unique_id(void *id128);
int  clbm_comm_init(clbm_ctx *ctx, const void *id128, int rank, int nranks);
int  clbm_slab_step(clbm_ctx *ctx, int nsteps);
int  clbm_comm_destroy(clbm_ctx *ctx);

/* ---- the ring over CUDA peer memory (the GPUs of one node: NVLink / NVSwitch) -- the default transport -------------
 * All halo blocks of a context live in one device allocation (the "mailbox").  clbm_peer_export writes an opaque
 * CLBM_PEER_HANDLE_BYTES-byte handle of it (a cudaIpcMemHandle plus the layout); the caller moves the handles of the two
 * ring neighbours to every rank by any means (torch.distributed all_gather, MPI, a file) and calls clbm_peer_connect
 * (left = the rank owning x_offset - 1, right = the rank owning x_offset + nx, periodic; a ring of two passes the same
 * handle twice).  From then on the pack of a halo phase writes straight into the neighbour's receive block and an exchange
 * is a flag store + a flag wait on the device; clbm_slab_step replays two captured steps per CUDA graph launch
 * (CLBM_SLAB_GRAPH=0 turns the capture off).  All ranks must pass a barrier between clbm_peer_connect and the first
 * clbm_slab_step / clbm_slab_exchange.  A wait that sees no signal for CLBM_PEER_TIMEOUT_MS (default 20000) gives up and
 * the next clbm_sync reports CLBM_ESTATE instead of hanging the GPU.
 * clbm_peer_connect_local is the same for contexts living in ONE process (tests; several GPUs driven by one process). */
#define CLBM_PEER_HANDLE_BYTES 128
int  clbm_peer_export(clbm_ctx *ctx, void *handle);
int  clbm_peer_connect(clbm_ctx *ctx, const void *left_handle, const void *right_handle);
int  clbm_peer_connect_local(clbm_ctx *ctx, clbm_ctx *left, clbm_ctx *right);
int  clbm_peer_disconnect(clbm_ctx *ctx);
/* 0: no ring, 1: NCCL ring (clbm_comm_init), 2: peer-memory ring */
int  clbm_ring_kind(const clbm_ctx *ctx);
/* pack + exchange + unpack of ONE halo phase on the launching stream, through whichever ring the context has
 * (phase 2: the node mask after clbm_upload; phase 0 after stage 0: the moment ghosts a field download needs) */
int  clbm_slab_exchange(clbm_ctx *ctx, int phase);
/* the two halves of an exchange on a peer ring, for a caller that issues the stages itself: clbm_slab_signal after the stage
 * that packed `phase`, clbm_slab_wait before the stage that unpacks it; boundary != 0: on the boundary stream (overlap
 * protocol).  Contexts of one process that share a GPU must issue all signals of a phase before any wait of it. */
int  clbm_slab_signal(clbm_ctx *ctx, int phase, int boundary);
int  clbm_slab_wait(clbm_ctx *ctx, int phase, int boundary);

/* ---- compliant-vessel case (CLBM_MODEL_PULSATILE) ---------------------------------------------------
 * Replaces the whole iteration body of PulsatileBloodFlow2D() ("Abbashub LBM/apps/PulsatileBloodFlow2D.h":764-790):
 *     for_each(par_unseq, lattice, lattice + nelem, lbm);   // MRT_Collision           :533-541, :672-676
 *     lbm.Boundary_Conditions();                            // Bouzidi_quadratic x2     :543-601
 *     lbm.Streaming();                                      // in-place pull            :603-616
 *     lbm.Inlet_ZouHe(t); lbm.Outlet_ZouHe(t);              //                          :618-669
 *     lbm.Macroscopic_Properties_g();                       //                          :216-230
 *     if (deformable) lbm.Calculate_Pressure_and_Move_Walls(t);  // + Fobj / border / fresh nodes  :233-498
 *     *parity = 1 - *parity;
 * The functor there owns P, Ux, Uy, yr1, yr2, Fobj and the border lists besides lattice/flag/parity, so this
 * model has its own context type.  Host arrays use the reference layout: lattice[p*9*nelem + k*nelem + i],
 * i = y + ny*x, nx = 1 + 10*(N-2), ny = N; flag uint8 {0 bounce_back, 1 bulk}.  Results are bit-identical to the
 * reference's (the device code is built without FMA contraction).                                            */
typedef struct clbm_pulsatile clbm_pulsatile;
typedef struct clbm_pulsatile_params {
    int32_t abi_version;    /* CLBM_ABI_VERSION */
    int32_t N;              /* ny = N, nx = 1 + 10*(N-2)  (AB:721-722) */
    int32_t device;         /* CUDA device ordinal, -1 = current */
    int32_t is_severed;     /* AB:105, :153-155, :741 */
    int32_t deformable;     /* AB:104, :740 */
    int32_t t_beat;         /* <= 0: max(1, nx)  (AB:749) */
    double tau;             /* s8 = 1/tau, s5 = 1  (AB:99-102, :742-744) */
    double alpha;           /* wall compliance  (AB:106, :745) */
    double p0_in, p0_out;   /* AB:746-747 (both 0: 0.20 / 0.19; severed: 0.02 / 0) */
} clbm_pulsatile_params;

/* Setup_Simulation_Parameters, Initialize_Yr_and_Vw_and_p, Initialize_Fobj_for_Vessel_Walls,
 * Find_or_Update_Boundary_Nodes, Initialize_P_U_g (AB:751-757), state resident on the device.
 * CLBM_EINVAL with "Initial wall location out of bounds." where the reference throws runtime_error (AB:181). */
int  clbm_pulsatile_create(const clbm_pulsatile_params *params, clbm_pulsatile **out);
int  clbm_pulsatile_destroy(clbm_pulsatile *ctx);
/* nx, ny, tf = t_beat + 2*t_propagation (AB:759), iterations done, current parity; any pointer may be NULL */
int  clbm_pulsatile_info(const clbm_pulsatile *ctx, int *nx, int *ny, int *tf, int *t_iter, int *parity);
/* nsteps iterations of the loop body above (asynchronous; sync/download report device-side failures) */
int  clbm_pulsatile_step(clbm_pulsatile *ctx, int nsteps);
int  clbm_pulsatile_step_timed(clbm_pulsatile *ctx, int nsteps, float *ms);
int  clbm_pulsatile_sync(clbm_pulsatile *ctx);
int64_t clbm_pulsatile_launch_count(const clbm_pulsatile *ctx);
/* device time per iteration (event pair around the kernels of each of the first cap_steps iterations) */
int  clbm_pulsatile_kernel_timing_begin(clbm_pulsatile *ctx, int cap_steps);
int  clbm_pulsatile_kernel_timing_end(clbm_pulsatile *ctx, float *avg_ms, int *count);
/* the functor's stored fields (what saveVtkFields_PulsatileBloodFlow2D prints, AB:680-706) and the wall positions */
int  clbm_pulsatile_download_fields(clbm_pulsatile *ctx, double *P, double *Ux, double *Uy, uint8_t *flag,
                                    double *yr1, double *yr2);
/* both lattice buffers (2*9*nelem doubles) and the parity */
int  clbm_pulsatile_download_lattice(clbm_pulsatile *ctx, double *lattice, int *parity);
/* hand over a host state built by the reference's own set-up code (both buffers, all functor fields) */
int  clbm_pulsatile_upload(clbm_pulsatile *ctx, const double *lattice, const uint8_t *flag, const double *P,
                           const double *Ux, const double *Uy, const double *yr1, const double *yr2,
                           int parity, int t_iter);

/* ---- Young-Laplace case (conservative phase field, AB/apps/Young_Laplace2D.h) ---------------------------------
 * Replaces the iteration body of Young_Laplace2D() ("Abbashub LBM/apps/Young_Laplace2D.h":555-565):
 *     for_each(par_unseq, cell_index.begin(), cell_index.end(), [&](int i){ lbm.collide_stream_at(i); });   // :217-290
 *     *parity = 1 - *parity;
 *     lbm.update_fields();                                                                                    // :297-370
 * Host arrays use the reference layout: lattice[h_in | h_out | g_in | g_out] (4*9*nelem doubles, in/out selected by
 * parity, :104-107), i = y + ny*x; fully periodic.  The functor's ten stored fields are not kept on the device; the
 * ones a driver prints (C, P, Rho, Ux, Uy) are evaluated on download with update_fields' definitions.        */
typedef struct clbm_yl2d clbm_yl2d;
typedef struct clbm_yl2d_params {
    int32_t abi_version;    /* CLBM_ABI_VERSION */
    int32_t nx, ny;         /* the driver uses N x N (AB:497) */
    int32_t device;         /* CUDA device ordinal, -1 = current */
    double Sigma, W, M;     /* surface tension, interface thickness, mobility (AB:85-87) */
    double RhoL, RhoH, tau; /* light / heavy density, hydrodynamic relaxation time (AB:83-84, :88) */
} clbm_yl2d_params;
/* parameters + iniCell on every node (AB:512-521); the first step performs the driver's initial update_fields */
int  clbm_yl2d_create(const clbm_yl2d_params *params, clbm_yl2d **out);
int  clbm_yl2d_destroy(clbm_yl2d *ctx);
int  clbm_yl2d_step(clbm_yl2d *ctx, int nsteps);
int  clbm_yl2d_step_timed(clbm_yl2d *ctx, int nsteps, float *ms);
int  clbm_yl2d_sync(clbm_yl2d *ctx);
int64_t clbm_yl2d_launch_count(const clbm_yl2d *ctx);
/* C (phi), P (p*), Rho, Ux, Uy as update_fields leaves them (what saveVtkFields_Young_Laplace2D prints, AB:374-421) */
int  clbm_yl2d_download_fields(clbm_yl2d *ctx, double *C, double *P, double *Rho, double *Ux, double *Uy);
int  clbm_yl2d_download_lattice(clbm_yl2d *ctx, double *lattice, int *parity);
/* hand over a state built by the reference's own code: populations + the velocity update_fields computed for them */
int  clbm_yl2d_upload(clbm_yl2d *ctx, const double *lattice, const double *Ux, const double *Uy, int parity);
/* CLBM_REDUCE_MASS: totalMass_Young_Laplace2D (AB:436-445); CLBM_REDUCE_ENERGY: computeEnergy_Young_Laplace2D (:425-435) */
int  clbm_yl2d_reduce(clbm_yl2d *ctx, int kind, double *out);

#ifdef __cplusplus
}
#endif
#endif /* CLBM_H */
