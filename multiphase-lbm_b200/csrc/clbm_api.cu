// clbm_api.cu -- the C ABI of include/clbm.h: context life cycle, state transfer between the
// reference host layout and the device slab storage, the step loop, diagnostics, profiling.
#include <sched.h>

#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <new>
#include <vector>

#include "clbm_internal.h"

namespace clbm {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return e == cudaErrorMemoryAllocation ? CLBM_ENOMEM : CLBM_ECUDA;
}

LaunchScope::LaunchScope(clbm_ctx *ctx, const char *n, bool dominant) : c(ctx), name(n), a(nullptr), b(nullptr), timed(false)
{
    c->launches++;
    if (dominant && c->ktiming && (int)c->kev.size() < 2 * c->ktiming_cap) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        c->kev.push_back(e0);
        c->kev.push_back(e1);
        cudaEventRecord(e0, c->stream);
        c->kname = n;
        timed = true;
    }
    if (c->profiling) {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, c->stream);
    }
}
LaunchScope::~LaunchScope()
{
    if (timed) cudaEventRecord(c->kev.back(), c->stream);
    if (c->profiling) {
        cudaEventRecord(b, c->stream);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        c->prof.push_back({name, ms});
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
}

// per-model dispatch (kernels live in the per-model .cu files)
int sc_fields(clbm_ctx *c, double *s0, double *s1, double *ux, double *uy, double *uz);
int sc_force_field(clbm_ctx *c, double *fx, double *fy, double *fz);
int diag_contact_angle(clbm_ctx *c, double rho_cut, int *base_y, int *base, int *height);          // diag_kernels.cu
int diag_contact_angle_raw(clbm_ctx *c, double rho_cut, int base_y_in, int out[4]);
int diag_interface_heights(clbm_ctx *c, double phi_mid, int *y_x0, int *y_xmid);
int hcz2d_fields(clbm_ctx *c, double *s0, double *s1, double *s2, double *ux, double *uy, double *uz);
int hcz3d_fields(clbm_ctx *c, double *s0, double *s1, double *s2, double *ux, double *uy, double *uz);
int slab_step_for_profile(clbm_ctx *c);   // slab_comm.cu
int sc_fused_multi_step(clbm_ctx *c, int nsteps, int *done);   // sc_fused.cu
int hcz2d_fused_multi_step(clbm_ctx *c, int nsteps, int *done);   // hcz2d_fused.cu
int sc_psi_all(clbm_ctx *c);
int sc_psi_boundary(clbm_ctx *c);
bool sc_range_supported(const clbm_ctx *c);
int sc_collide_range_fused(clbm_ctx *c, int x_begin, int x_end, int x2_begin, int x2_end);
int sc_collide_slab(clbm_ctx *c);
int hcz2d_stage0(clbm_ctx *c);
int hcz2d_stage1(clbm_ctx *c);
bool hcz2d_fused_eligible(const clbm_ctx *c);                      // hcz2d_fused.cu
int hcz2d_fused_range(clbm_ctx *c, int x_begin, int x_end);
int hcz3d_stage0(clbm_ctx *c, bool rebuild);
int hcz3d_stage1(clbm_ctx *c);

static int model_step(clbm_ctx *c)
{
    switch (c->prm.model) {
    case CLBM_MODEL_SC_D2Q9:
    case CLBM_MODEL_SC_D3Q19: return sc_step(c);
    case CLBM_MODEL_HCZ_D2Q9: return hcz2d_step(c);
    case CLBM_MODEL_HCZ_D3Q19: return hcz3d_step(c);
    }
    set_error("model %d has no step", c->prm.model);
    return CLBM_EINVAL;
}

int model_fields(clbm_ctx *c, double *s0, double *s1, double *s2, double *ux, double *uy, double *uz)
{
    switch (c->prm.model) {
    case CLBM_MODEL_SC_D2Q9:
    case CLBM_MODEL_SC_D3Q19: return sc_fields(c, s0, s1, ux, uy, uz);
    case CLBM_MODEL_HCZ_D2Q9: return hcz2d_fields(c, s0, s1, s2, ux, uy, uz);
    case CLBM_MODEL_HCZ_D3Q19: return hcz3d_fields(c, s0, s1, s2, ux, uy, uz);
    }
    return CLBM_EINVAL;
}

// The overlap protocol needs x-range launches of the collide kernel: the TMA Shan-Chen kernel and the fused HCZ D2Q9 kernel
// have them.  D = number of boundary planes per side whose stencils reach into the neighbour slab (= the halo depth of the
// moment exchange: psi depth 1, phi depth 2).
// Default per model, from the measurements of DESIGN.md section 4 (tools/self_ring_bench.py, peer ring, graph replay):
//   Shan-Chen (TMA kernel): the SEQUENTIAL protocol.  64-plane slab 1225 us per step against 1243 (interior first) / 1226 (halo
//   first); 512-plane slab 8377 / 8412 / 8388.  With the peer ring an exchange is ~15 us; the boundary-plane launches of the
//   overlap forms cost the full-SM kernel more than that (the interior launch: 1007 us alone, 1220 us with them in its waves).
//   HCZ D2Q9: interior first.  256-column slab 162 us against 174 (sequential) / 184 (halo first); 2048 columns 1039 / 1058 / 1076.
#define OVERLAP_DEFAULT(c) ((c)->prm.model == CLBM_MODEL_HCZ_D2Q9 ? 1 : 0)

static bool overlap_possible(const clbm_ctx *c)
{
    const int m = c->prm.model;
    if (!c->multi) return false;
    if (m == CLBM_MODEL_SC_D2Q9 || m == CLBM_MODEL_SC_D3Q19) return c->geo.nx >= 3 && sc_range_supported(c);
    if (m == CLBM_MODEL_HCZ_D2Q9) return c->geo.nx >= 4 && c->prm.fused && hcz2d_fused_eligible(c);
    return false;
}
// Which form of the overlap protocol a slab step uses (0: none, the sequential stages 0-2).  Two forms exist because neither
// wins everywhere (tools/self_ring_bench.py):
//   1  interior first: the interior planes are collided on the launching stream WHILE the boundary stream exchanges the moment
//      halo, collides the boundary planes and exchanges the crossing populations;
//   2  halo first: moment halo exchanged on the launching stream, a chunk of boundary planes per side collided, then the
//      interior, which overlaps with the exchange of the crossing populations only.
// CLBM_SLAB_OVERLAP = 0 / 1 / 2 forces one (read in clbm_create).
static int overlap_variant(const clbm_ctx *c)
{
    if (!overlap_possible(c)) return 0;
    if (c->env.slab_overlap >= 0) return c->env.slab_overlap > 2 ? 2 : c->env.slab_overlap;
    return OVERLAP_DEFAULT(c);
}
static bool overlap_supported(const clbm_ctx *c) { return overlap_variant(c) != 0; }
static int overlap_depth(const clbm_ctx *c) { return c->prm.model == CLBM_MODEL_HCZ_D2Q9 ? 2 : 1; }
static int overlap_moments(clbm_ctx *c) { return c->prm.model == CLBM_MODEL_HCZ_D2Q9 ? hcz2d_stage0(c) : sc_psi_boundary(c); }
// collide + push of [x_begin, x_end) and, when non-empty, [x2_begin, x2_end)
static int overlap_collide(clbm_ctx *c, int x_begin, int x_end, int x2_begin, int x2_end)
{
    if (c->prm.model != CLBM_MODEL_HCZ_D2Q9) return sc_collide_range_fused(c, x_begin, x_end, x2_begin, x2_end);
    int rc = hcz2d_fused_range(c, x_begin, x_end);
    if (rc || x2_end <= x2_begin) return rc;
    return hcz2d_fused_range(c, x2_begin, x2_end);
}

static int ensure_boundary_stream(clbm_ctx *c)
{
    if (c->stream_b) return 0;
    int lo = 0, hi = 0;
    CLBM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CLBM_CUDA(cudaStreamCreateWithPriority(&c->stream_b, cudaStreamNonBlocking, hi));
    CLBM_CUDA(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    CLBM_CUDA(cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming));
    CLBM_CUDA(cudaEventRecord(c->ev_main, c->stream));
    return 0;
}

// while alive, the boundary stream is the context's launching stream
struct BoundaryStream {
    clbm_ctx *c;
    cudaStream_t saved;
    explicit BoundaryStream(clbm_ctx *ctx) : c(ctx), saved(ctx->stream) { c->stream = c->stream_b; }
    ~BoundaryStream() { c->stream = saved; }
};

// Overlap protocol of one slab step (stages 10, 11, 12; include/clbm.h), second form (round 2).
//
// Round 1 collided the interior planes WHILE the boundary stream ran the moment exchange and the two boundary planes.  With the
// peer-memory ring the moment exchange costs ~15 us, and the measurement on a 64-plane slab (tools/self_ring_bench.py) showed
// what the concurrency cost: the interior launch takes 1007 us alone and 1220 us with the high-priority boundary-plane CTAs of
// the same kernel cutting into its waves (they break the L2 sharing of neighbouring planes that the short x-chunks live on).
// So the moment halo is now exchanged FIRST, on the launching stream (stage 10 + exchange 0), then the boundary planes -- a
// chunk of Bw planes on each side, not one plane, so that their two-plane prologue is amortised -- are collided, and only
// then the interior; what overlaps with the interior is the exchange of the crossing populations (pack, signal, wait,
// unpack on the boundary stream), which is the larger message and the one whose latency would otherwise sit between two steps.
static int overlap_width(const clbm_ctx *c)
{
    const int nx = c->geo.nx, D = overlap_depth(c);
    int bw = c->prm.model == CLBM_MODEL_HCZ_D2Q9 ? 16 : 8;
    if (bw > nx / 4) bw = nx / 4;
    return bw < D ? D : bw;
}

// form 1 (round 1): see overlap_variant
static int overlap_stage_interior_first(clbm_ctx *c, int stage)
{
    if (!overlap_supported(c)) { set_error("overlap protocol not available for this context (use stages 0-2)"); return CLBM_ESTATE; }
    int rc;
    const int nx = c->geo.nx, D = overlap_depth(c);
    if ((rc = ensure_boundary_stream(c))) return rc;
    if (stage == 10) {
        {   // the boundary planes of the "in" buffer were completed by the launching stream: by the previous step's interior
            // launch, or -- after a step of the sequential protocol, an upload or a device-side init -- by whatever that stream
            // ran last.  Recording here (not only after the interior launch) orders ALL of it before the boundary stream.
            CLBM_CUDA(cudaEventRecord(c->ev_main, c->stream));
            CLBM_CUDA(cudaStreamWaitEvent(c->stream_b, c->ev_main, 0));
            BoundaryStream bs(c);
            if ((rc = overlap_moments(c))) return rc;
            if ((rc = halo_pack(c, 0))) return rc;
            // the interior launch fills every SM for the rest of the step: let these two small kernels through first
            // (they run in ~20 us on the idle GPU; behind the interior's first wave they took 350 us)
            CLBM_CUDA(cudaEventRecord(c->ev_b, c->stream_b));
        }
        CLBM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_b, 0));
        if ((rc = overlap_collide(c, D, nx - D, 0, 0))) return rc;
        CLBM_CUDA(cudaEventRecord(c->ev_main, c->stream));
        return 0;
    }
    if (stage == 11) {
        BoundaryStream bs(c);
        if ((rc = halo_unpack(c, 0))) return rc;
        if ((rc = overlap_collide(c, 0, D, nx - D, nx))) return rc;   // SC: both boundary planes in one launch
        c->parity = 1 - c->parity;
        return halo_pack(c, 1);
    }
    {
        BoundaryStream bs(c);
        if ((rc = halo_unpack(c, 1))) return rc;
    }
    CLBM_CUDA(cudaEventRecord(c->ev_b, c->stream_b));
    CLBM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_b, 0));
    return 0;
}


// form 2
static int overlap_stage_halo_first(clbm_ctx *c, int stage)
{
    if (!overlap_supported(c)) { set_error("overlap protocol not available for this context (use stages 0-2)"); return CLBM_ESTATE; }
    int rc;
    const int nx = c->geo.nx, Bw = overlap_width(c);
    if ((rc = ensure_boundary_stream(c))) return rc;
    if (stage == 10) {   // launching stream: moments of the boundary planes + pack of the moment halo (exchange 0 follows on this stream)
        if ((rc = overlap_moments(c))) return rc;
        return halo_pack(c, 0);
    }
    if (stage == 11) {
        if ((rc = halo_unpack(c, 0))) return rc;
        if (nx <= 2 * Bw) {
            if ((rc = overlap_collide(c, 0, nx, 0, 0))) return rc;              // a slab this thin has no interior worth a launch
            CLBM_CUDA(cudaEventRecord(c->ev_main, c->stream));
        } else {
            if ((rc = overlap_collide(c, 0, Bw, nx - Bw, nx))) return rc;       // both boundary chunks in one launch
            CLBM_CUDA(cudaEventRecord(c->ev_main, c->stream));                  // the ghost planes hold what crossed the faces
            if ((rc = overlap_collide(c, Bw, nx - Bw, 0, 0))) return rc;        // interior: overlaps with exchange 1
        }
        c->parity = 1 - c->parity;
        CLBM_CUDA(cudaStreamWaitEvent(c->stream_b, c->ev_main, 0));
        BoundaryStream bs(c);
        return halo_pack(c, 1);
    }
    {
        BoundaryStream bs(c);
        if ((rc = halo_unpack(c, 1))) return rc;
    }
    CLBM_CUDA(cudaEventRecord(c->ev_b, c->stream_b));
    CLBM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_b, 0));
    return 0;
}

static int overlap_stage(clbm_ctx *c, int stage)
{
    return overlap_variant(c) == 2 ? overlap_stage_halo_first(c, stage) : overlap_stage_interior_first(c, stage);
}

// slab protocol: see include/clbm.h (clbm_step_stage)
int model_stage(clbm_ctx *c, int stage)
{
    int rc = 0;
    const int m = c->prm.model;
    if (stage == 0 || stage == 20) {   // 20: stage 0 with every moment rebuilt from the populations (before a field download)
        if (m == CLBM_MODEL_SC_D2Q9 || m == CLBM_MODEL_SC_D3Q19) rc = c->prm.fused ? sc_psi_boundary(c) : sc_psi_all(c);
        else if (m == CLBM_MODEL_HCZ_D2Q9) rc = hcz2d_stage0(c);
        else if (m == CLBM_MODEL_HCZ_D3Q19) rc = hcz3d_stage0(c, stage == 20);
        else rc = CLBM_EINVAL;
        if (rc) return rc;
        return halo_pack(c, 0);
    }
    if (stage == 1) {
        if ((rc = halo_unpack(c, 0))) return rc;
        if (m == CLBM_MODEL_SC_D2Q9 || m == CLBM_MODEL_SC_D3Q19) rc = sc_collide_slab(c);
        else if (m == CLBM_MODEL_HCZ_D2Q9) rc = hcz2d_stage1(c);
        else if (m == CLBM_MODEL_HCZ_D3Q19) rc = hcz3d_stage1(c);
        else rc = CLBM_EINVAL;
        if (rc) return rc;
        c->parity = 1 - c->parity;   // the freshly written buffer becomes "in"
        return halo_pack(c, 1);
    }
    if (stage == 2) return halo_unpack(c, 1);
    if (stage >= 10 && stage <= 12) return overlap_stage(c, stage);
    set_error("bad stage %d", stage);
    return CLBM_EINVAL;
}

// device scratch of the field / force downloads: allocated once at the largest size asked for and kept with the context
// (a cudaMalloc + cudaFree pair per call is a device-wide synchronisation and milliseconds at 512^3)
int field_scratch(clbm_ctx *c, size_t bytes, double **out)
{
    if (c->scratch_bytes < bytes) {
        if (c->scratch) { CLBM_CUDA(cudaStreamSynchronize(c->stream)); cudaFree(c->scratch); c->scratch = nullptr; c->scratch_bytes = 0; }
        if (cudaMalloc(&c->scratch, bytes) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory (%zu bytes of field scratch)", bytes); return CLBM_ENOMEM; }
        c->scratch_bytes = bytes;
    }
    *out = c->scratch;
    return 0;
}

// while alive, the calling thread is restricted to the CPUs of the NUMA node the current CUDA device hangs off
struct NumaPin {
    cpu_set_t old;
    bool active = false;
    NumaPin()
    {
        int dev = 0;
        char bus[32] = "", path[128], buf[4096];
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) { cudaGetLastError(); return; }
        for (char *q = bus; *q; ++q) *q = (char)tolower(*q);
        snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
        FILE *f = fopen(path, "r");
        int node = -1;
        if (f) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
        if (node < 0) return;
        snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
        f = fopen(path, "r");
        if (!f) return;
        const bool got = fgets(buf, sizeof(buf), f) != nullptr;
        fclose(f);
        if (!got) return;
        cpu_set_t want, allowed;
        CPU_ZERO(&want);
        for (char *tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {   // "0-15,32-47"
            int a = 0, b = 0;
            const int n = sscanf(tok, "%d-%d", &a, &b);
            if (n == 1) b = a;
            if (n >= 1) for (int i = a; i <= b && i < CPU_SETSIZE; ++i) CPU_SET(i, &want);
        }
        if (sched_getaffinity(0, sizeof(old), &old) != 0) return;
        CPU_AND(&allowed, &want, &old);
        if (CPU_COUNT(&allowed) == 0) return;
        active = sched_setaffinity(0, sizeof(allowed), &allowed) == 0;
    }
    void restore() { if (active) { sched_setaffinity(0, sizeof(old), &old); active = false; } }
    ~NumaPin() { restore(); }
};

}  // namespace clbm

using namespace clbm;

extern "C" {

const char *clbm_last_error(void) { return g_err; }
int clbm_abi_version(void) { return CLBM_ABI_VERSION; }

int clbm_create(const clbm_params *p, clbm_ctx **out)
{
    if (!p || !out) { set_error("null argument"); return CLBM_EINVAL; }
    *out = nullptr;
    if (p->abi_version != CLBM_ABI_VERSION) { set_error("ABI version %d != %d", p->abi_version, CLBM_ABI_VERSION); return CLBM_EINVAL; }
    if (p->model < 0 || p->model > CLBM_MODEL_HCZ_D3Q19) { set_error("unsupported model %d", p->model); return CLBM_EINVAL; }
    if (p->sc_force == CLBM_SC_FORCE_CONSTG && p->G == 0.0) { set_error("constant-G Shan-Chen needs G != 0"); return CLBM_EINVAL; }
    if ((p->model == CLBM_MODEL_SC_D2Q9 || p->model == CLBM_MODEL_SC_D3Q19) && (p->sc_force < 0 || p->sc_force > CLBM_SC_FORCE_EXPGUO)) {
        set_error("unknown Shan-Chen force variant %d", p->sc_force);
        return CLBM_EINVAL;
    }
    if (p->collision != CLBM_COLLISION_BGK) {
        if (p->collision != CLBM_COLLISION_MRT) {
            set_error("collision operator %d: BGK (0) and MRT (1) exist", p->collision);
            return CLBM_EINVAL;
        }
        const double r[3] = {p->s_e, p->s_eps, p->s_q};
        for (double v : r)
            if (!(v > 0.0 && v < 2.0)) { set_error("MRT rates s_e, s_eps, s_q must lie in (0, 2): got %g %g %g", r[0], r[1], r[2]); return CLBM_EINVAL; }
    }
    if (p->model == CLBM_MODEL_SC_D3Q19 && p->sc_force == CLBM_SC_FORCE_EXPGUO) {
        set_error("the psi = 1 - exp(-rho) / Guo variant (SC/apps/RayleighTaylor2D.h) is D2Q9 only");
        return CLBM_EINVAL;
    }
    const bool is3d = p->model == CLBM_MODEL_SC_D3Q19 || p->model == CLBM_MODEL_HCZ_D3Q19;
    if (p->nx < 1 || p->ny < 1 || p->nz < 1 || (!is3d && p->nz != 1)) { set_error("bad extent %d x %d x %d", p->nx, p->ny, p->nz); return CLBM_EINVAL; }
    if (p->nx_global < p->nx || p->x_offset < 0 || p->x_offset + p->nx > p->nx_global) { set_error("bad slab [%d,%d) of %d", p->x_offset, p->x_offset + p->nx, p->nx_global); return CLBM_EINVAL; }
    if (!(p->omega > 0.0)) { set_error("omega must be > 0"); return CLBM_EINVAL; }

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: this library has no CPU fallback");
        return CLBM_ENODEVICE;
    }
    int dev = p->device;
    if (dev < 0) CLBM_CUDA(cudaGetDevice(&dev));
    CLBM_CUDA(cudaSetDevice(dev));

    clbm_ctx *c = new (std::nothrow) clbm_ctx();
    if (!c) return CLBM_ENOMEM;
    memset((void *)&c->prm, 0, sizeof(c->prm));
    c->prm = *p;
    // the MRT operator of the HCZ D3Q19 model lives in the staged collide kernel only
    if (p->model == CLBM_MODEL_HCZ_D3Q19 && p->collision == CLBM_COLLISION_MRT) c->prm.fused = 0;
    c->device = dev;
    c->Q = is3d ? 19 : 9;
    c->sets = (p->model == CLBM_MODEL_HCZ_D2Q9 || p->model == CLBM_MODEL_HCZ_D3Q19) ? 2 : 1;
    c->multi = p->nx != p->nx_global || env_int("CLBM_FORCE_SLAB") == 1;
    c->parity = 0;
    c->host_parity0 = 0;
    c->steps_taken = 0;
    c->launches = 0;
    c->profiling = false;
    c->ktiming = false;
    c->ktiming_cap = 0;
    c->kname = "";
    c->stage = nullptr;
    c->stage_bytes = 0;
    c->mailbox = nullptr;
    c->mailbox_bytes = c->mailbox_flags_off = 0;
    c->peer_mode = 0;
    c->peer_base[0] = c->peer_base[1] = nullptr;
    c->peer_err = nullptr;
    c->ring_fuse = 0;
    c->halo0_packed = 0;
    c->fld0_in_mailbox = 0;
    c->mailbox_psi_off = 0;
    c->halo0_direct = 0;
    c->peer_nx[0] = c->peer_nx[1] = 0;
    c->slab_graph[0] = c->slab_graph[1] = nullptr;
    c->slab_graph_launches[0] = c->slab_graph_launches[1] = 0;
    c->slab_graph_failed = 0;
    c->comm = nullptr;
    c->comm_rank = 0;
    c->comm_size = 0;
    c->nfld = 0;
    c->scratch = nullptr;
    c->scratch_bytes = 0;
    c->sc_queue = nullptr;
    c->resident_progress = nullptr;
    c->resident_epoch = 0;
    c->sc_queue_next = 0;
    c->sm_count = 0;
    c->stream_u = nullptr;
    memset(c->mom, 0, sizeof(c->mom));
    memset(c->mome, 0, sizeof(c->mome));
    c->mom_src = c->mom_valid = c->sweep_active = 0;
    c->walls_known = c->has_walls = 0;
    read_env_knobs(c->env);
    for (auto &s : c->pop) for (auto &b : s) b = nullptr;
    for (auto &f : c->fld) f = nullptr;
    memset(c->halo, 0, sizeof(c->halo));
    memset(c->halo_bytes, 0, sizeof(c->halo_bytes));
    c->flag = nullptr; c->red_dev = nullptr; c->red_host = nullptr;

    Geom &g = c->geo;
    g.nx = p->nx; g.ny = p->ny; g.nz = p->nz;
    g.G = (p->model == CLBM_MODEL_HCZ_D3Q19) ? 3 : (p->model == CLBM_MODEL_HCZ_D2Q9 ? 2 : 1);
    g.wrapx = c->multi ? 0 : 1;
    g.nx_global = p->nx_global; g.x_offset = p->x_offset;
    g.plane = (long long)p->ny * p->nz;
    g.ncs = (long long)(p->nx + 2 * g.G) * g.plane;
    if (c->multi && p->nx < g.G) { set_error("slab thinner (%d) than the halo depth (%d)", p->nx, g.G); delete c; return CLBM_EINVAL; }

    derive_model_params(p, c->mp);

    int rc = 0;
    auto fail = [&](int code) { clbm_destroy(c); return code; };
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream create failed"); delete c; return CLBM_ECUDA; }
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    const size_t popbytes = (size_t)c->Q * g.ncs * sizeof(double);
    for (int s = 0; s < c->sets; ++s)
        for (int b = 0; b < 2; ++b) {
            if (cudaMalloc(&c->pop[s][b], popbytes) != cudaSuccess) { set_error("out of device memory allocating %zu bytes of populations", popbytes); return fail(CLBM_ENOMEM); }
            cudaMemsetAsync(c->pop[s][b], 0, popbytes, c->stream);
        }
    if (cudaMalloc(&c->flag, (size_t)g.ncs) != cudaSuccess) { set_error("out of device memory (flag)"); return fail(CLBM_ENOMEM); }
    cudaMemsetAsync(c->flag, CELL_BULK, (size_t)g.ncs, c->stream);
    c->nfld = (p->model == CLBM_MODEL_HCZ_D3Q19) ? 8 : (p->model == CLBM_MODEL_HCZ_D2Q9 ? 5 : 1);
    // (the psi field of a Shan-Chen x-slab is placed inside the halo mailbox by halo_alloc below)
    c->fld0_in_mailbox = c->multi && (p->model == CLBM_MODEL_SC_D2Q9 || p->model == CLBM_MODEL_SC_D3Q19);
    for (int i = c->fld0_in_mailbox ? 1 : 0; i < c->nfld; ++i) {
        if (cudaMalloc(&c->fld[i], (size_t)g.ncs * sizeof(double)) != cudaSuccess) { set_error("out of device memory (field %d)", i); return fail(CLBM_ENOMEM); }
        cudaMemsetAsync(c->fld[i], 0, (size_t)g.ncs * sizeof(double), c->stream);
    }
    if (cudaMalloc(&c->red_dev, 8 * sizeof(double)) != cudaSuccess || cudaMallocHost(&c->red_host, 8 * sizeof(double)) != cudaSuccess) {
        set_error("out of memory (reduction scratch)");
        return fail(CLBM_ENOMEM);
    }
    if (c->multi && (rc = halo_alloc(c))) return fail(rc);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { set_error("device initialisation failed"); return fail(CLBM_ECUDA); }
    *out = c;
    return CLBM_OK;
}

int clbm_destroy(clbm_ctx *c)
{
    if (!c) return CLBM_OK;
    cudaSetDevice(c->device);
    clbm_comm_destroy(c);
    clbm_peer_disconnect(c);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto &s : c->pop) for (auto &b : s) if (b) cudaFree(b);
    if (c->fld0_in_mailbox) c->fld[0] = nullptr;   // part of the mailbox allocation
    for (auto &f : c->fld) if (f) cudaFree(f);
    if (c->flag) cudaFree(c->flag);
    if (c->red_dev) cudaFree(c->red_dev);
    if (c->red_host) cudaFreeHost(c->red_host);
    if (c->stage) cudaFreeHost(c->stage);
    if (c->scratch) cudaFree(c->scratch);
    if (c->sc_queue) cudaFree(c->sc_queue);
    if (c->resident_progress) cudaFree(c->resident_progress);
    for (int m = 0; m < 5; ++m) {
        if (c->mom[1][m]) cudaFree(c->mom[1][m]);      // mom[0] aliases fld[0..4]
        for (int s = 0; s < 2; ++s) if (c->mome[s][m]) cudaFree(c->mome[s][m]);
    }
    if (c->mailbox) cudaFree(c->mailbox);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream_b) {
        cudaStreamSynchronize(c->stream_b);
        cudaStreamDestroy(c->stream_b);
        cudaEventDestroy(c->ev_main);
        cudaEventDestroy(c->ev_b);
    }
    if (c->stream_u) cudaStreamDestroy(c->stream_u);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return CLBM_OK;
}

int clbm_upload(clbm_ctx *c, const double *lattice, const uint8_t *flag, int parity) { return clbm_upload2(c, lattice, flag, parity, 1); }

int clbm_upload2(clbm_ctx *c, const double *lattice, const uint8_t *flag, int parity, int other_buffer)
{
    if (!c || !lattice || !flag || (parity != 0 && parity != 1)) { set_error("bad argument to clbm_upload"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    const Geom &g = c->geo;
    const size_t nelem = (size_t)g.nx * g.plane, npop = (size_t)c->Q * nelem;
    const size_t ghost = (size_t)g.G * g.plane;
    // the device "in" buffer becomes buffer 0; the other one restarts from zero like the reference's
    // value-initialised vector
    c->parity = 0;
    // the population arrays alternate between two streams so that two host-to-device copies are in flight at a time (one
    // stream keeps a single DMA engine busy and leaves the link idle between the 19 / 38 copies)
    if (!c->stream_u) CLBM_CUDA(cudaStreamCreateWithFlags(&c->stream_u, cudaStreamNonBlocking));
    int turn = 0;
    for (int s = 0; s < c->sets; ++s) {
        CLBM_CUDA(cudaMemsetAsync(c->pop[s][1], 0, (size_t)c->Q * g.ncs * sizeof(double), c->stream));
        const double *src = lattice + (size_t)s * 2 * npop + (size_t)parity * npop;
        for (int k = 0; k < c->Q; ++k)
            CLBM_CUDA(cudaMemcpyAsync(c->pop[s][0] + (size_t)k * g.ncs + ghost, src + (size_t)k * nelem, nelem * sizeof(double), cudaMemcpyHostToDevice,
                                      (turn++ & 1) ? c->stream_u : c->stream));
    }
    CLBM_CUDA(cudaMemcpyAsync(c->flag + ghost, flag, nelem, cudaMemcpyHostToDevice, c->stream));
    CLBM_CUDA(cudaStreamSynchronize(c->stream_u));
    // The buffer the caller did not select: every slot of a bulk node is rewritten by each step, so only bounce_back nodes
    // keep their initial values there.  The reference's inits leave them zero, except the layered HCZ one that fills both
    // buffers (PF/apps/twoLayeredFlow2D.h:184-187): hand those node values over as well, so a later
    // clbm_download_lattice returns what the reference's host array would hold at EVERY node.
    if (other_buffer) {
        std::vector<long long> widx;
        // CellType has two values: memchr finds the bounce_back (0) bytes at memory speed
        for (const uint8_t *q = flag, *end = flag + nelem; q < end && (q = (const uint8_t *)memchr(q, CELL_BB, (size_t)(end - q))); ++q)
            widx.push_back((long long)(ghost + (size_t)(q - flag)));
        const size_t nn = widx.size();
        if (nn) {
            std::vector<double> vals((size_t)c->sets * c->Q * nn);
            bool any = false;
            for (int s = 0; s < c->sets; ++s)
                for (int k = 0; k < c->Q; ++k) {
                    const double *src = lattice + (size_t)s * 2 * npop + (size_t)(1 - parity) * npop + (size_t)k * nelem;
                    double *dst = vals.data() + ((size_t)s * c->Q + k) * nn;
                    for (size_t j = 0; j < nn; ++j) { dst[j] = src[widx[j] - (long long)ghost]; any |= dst[j] != 0.0; }
                }
            if (any) {
                long long *idx_dev = nullptr;
                double *vals_dev = nullptr;
                if (cudaMalloc(&idx_dev, nn * sizeof(long long)) != cudaSuccess || cudaMalloc(&vals_dev, vals.size() * sizeof(double)) != cudaSuccess) {
                    cudaGetLastError();
                    if (idx_dev) cudaFree(idx_dev);
                    set_error("out of device memory (bounce_back node hand-over)");
                    return CLBM_ENOMEM;
                }
                int rc = 0;
                if (cudaMemcpyAsync(idx_dev, widx.data(), nn * sizeof(long long), cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
                    cudaMemcpyAsync(vals_dev, vals.data(), vals.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
                    rc = CLBM_ECUDA;
                if (!rc) rc = scatter_node_pops(c, 1, idx_dev, vals_dev, (long long)nn);
                cudaStreamSynchronize(c->stream);
                cudaFree(idx_dev);
                cudaFree(vals_dev);
                if (rc) { set_error("bounce_back node hand-over failed"); return rc; }
            }
        }
    }
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    c->host_parity0 = parity;   // download returns (uploaded parity + steps taken) & 1, like the reference's *parity
    c->steps_taken = 0;
    c->mom_valid = 0;
    c->walls_known = 0;
    return CLBM_OK;
}

static int host_parity(const clbm_ctx *c) { return (int)((c->host_parity0 + c->steps_taken) & 1); }
static void count_step(clbm_ctx *c) { c->steps_taken++; }

int clbm_download_lattice(clbm_ctx *c, double *lattice, int *parity)
{
    if (!c || !lattice) { set_error("bad argument to clbm_download_lattice"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    const Geom &g = c->geo;
    const size_t nelem = (size_t)g.nx * g.plane, npop = (size_t)c->Q * nelem;
    const size_t ghost = (size_t)g.G * g.plane;
    const int hp = host_parity(c);
    for (int s = 0; s < c->sets; ++s) {
        double *dst = lattice + (size_t)s * 2 * npop + (size_t)hp * npop;
        for (int k = 0; k < c->Q; ++k)
            CLBM_CUDA(cudaMemcpyAsync(dst + (size_t)k * nelem, c->pop[s][c->parity] + (size_t)k * g.ncs + ghost, nelem * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    if (parity) *parity = hp;
    return CLBM_OK;
}

int clbm_download_fields(clbm_ctx *c, double *s0, double *s1, double *s2, double *ux, double *uy, double *uz, uint8_t *flag)
{
    if (!c) { set_error("null context"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    const Geom &g = c->geo;
    const size_t nelem = (size_t)g.nx * g.plane;
    double *host[6] = {s0, s1, s2, ux, uy, uz};
    double *dev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int nwant = 0;
    for (auto h : host) nwant += h != nullptr;
    if (nwant) {
        double *tmp = nullptr;
        if (int rc = field_scratch(c, (size_t)nwant * nelem * sizeof(double), &tmp)) return rc;
        int j = 0;
        for (int i = 0; i < 6; ++i) if (host[i]) dev[i] = tmp + (size_t)(j++) * nelem;
        int rc = model_fields(c, dev[0], dev[1], dev[2], dev[3], dev[4], dev[5]);
        if (rc) return rc;
        for (int i = 0; i < 6; ++i)
            if (host[i]) CLBM_CUDA(cudaMemcpyAsync(host[i], dev[i], nelem * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (flag) CLBM_CUDA(cudaMemcpyAsync(flag, c->flag + (size_t)g.G * g.plane, nelem, cudaMemcpyDeviceToHost, c->stream));
        CLBM_CUDA(cudaStreamSynchronize(c->stream));
        return CLBM_OK;
    }
    if (flag) {
        CLBM_CUDA(cudaMemcpyAsync(flag, c->flag + (size_t)g.G * g.plane, nelem, cudaMemcpyDeviceToHost, c->stream));
        CLBM_CUDA(cudaStreamSynchronize(c->stream));
    }
    return CLBM_OK;
}

int clbm_download_force(clbm_ctx *c, double *fx, double *fy, double *fz)
{
    if (!c) { set_error("null context"); return CLBM_EINVAL; }
    if (c->prm.model != CLBM_MODEL_SC_D2Q9 && c->prm.model != CLBM_MODEL_SC_D3Q19) {
        set_error("clbm_download_force: Shan-Chen models only");
        return CLBM_EINVAL;
    }
    CLBM_CUDA(cudaSetDevice(c->device));
    const size_t nelem = (size_t)c->geo.nx * c->geo.plane;
    double *host[3] = {fx, fy, fz}, *dev[3] = {nullptr, nullptr, nullptr}, *tmp = nullptr;
    if (int rc = field_scratch(c, 3 * nelem * sizeof(double), &tmp)) return rc;
    for (int i = 0; i < 3; ++i) if (host[i]) dev[i] = tmp + (size_t)i * nelem;
    int rc = sc_force_field(c, dev[0], dev[1], dev[2]);
    for (int i = 0; i < 3 && !rc; ++i)
        if (host[i] && cudaMemcpyAsync(host[i], dev[i], nelem * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) {
            set_error("force download failed");
            rc = CLBM_ECUDA;
        }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (!rc && e != cudaSuccess) return cuda_fail(e, "force download sync", __FILE__, __LINE__);
    return rc;
}

int clbm_init_case(clbm_ctx *c, int case_id, const double *args, int nargs)
{
    if (!c) { set_error("null context"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    int rc = model_init_case(c, case_id, args, nargs);
    if (rc) return rc;
    c->host_parity0 = 0;
    c->steps_taken = 0;
    c->mom_valid = 0;
    c->walls_known = 0;
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    return CLBM_OK;
}

int clbm_step(clbm_ctx *c, int nsteps)
{
    if (!c || nsteps < 0) { set_error("bad argument to clbm_step"); return CLBM_EINVAL; }
    if (c->multi) { set_error("clbm_step on an x-slab: drive it with clbm_step_stage + halo exchange"); return CLBM_ESTATE; }
    CLBM_CUDA(cudaSetDevice(c->device));
    if (c->prm.model == CLBM_MODEL_SC_D2Q9 && nsteps >= 2) {
        // L2-resident D2Q9 lattices: all the steps in one cooperative launch (sc_fused.cu)
        int done = 0;
        if (int rc = sc_fused_multi_step(c, nsteps, &done)) return rc;
        if (done) { c->steps_taken += nsteps; return CLBM_OK; }
    }
    if (c->prm.model == CLBM_MODEL_HCZ_D2Q9 && nsteps >= 2) {
        // HCZ D2Q9, opt-in (CLBM_HCZ2D_MULTI=1; hcz2d_fused.cu, MULTI form)
        int done = 0;
        if (int rc = hcz2d_fused_multi_step(c, nsteps, &done)) return rc;
        if (done) { c->steps_taken += nsteps; return CLBM_OK; }
    }
    for (int s = 0; s < nsteps; ++s) {
        int rc = model_step(c);
        if (rc) return rc;
        count_step(c);
    }
    return CLBM_OK;
}

int clbm_step_stage(clbm_ctx *c, int stage)
{
    if (!c) { set_error("null context"); return CLBM_EINVAL; }
    if (!c->multi) { set_error("clbm_step_stage needs an x-slab context (nx < nx_global)"); return CLBM_ESTATE; }
    CLBM_CUDA(cudaSetDevice(c->device));
    int rc = model_stage(c, stage);
    if (rc) return rc;
    if (stage == 1 || stage == 11) count_step(c);
    return CLBM_OK;
}

int clbm_overlap_supported(const clbm_ctx *c) { return c && overlap_supported(c) ? 1 : 0; }
int clbm_overlap_variant(const clbm_ctx *c) { return c ? overlap_variant(c) : 0; }
int clbm_overlap_width(const clbm_ctx *c)
{
    if (!c) return 0;
    const int v = overlap_variant(c);
    if (v == 1) return overlap_depth(c);
    if (v == 2) return c->geo.nx <= 2 * overlap_width(c) ? 0 : overlap_width(c);
    return 0;
}

int clbm_sync(clbm_ctx *c)
{
    if (!c) { set_error("null context"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    if (c->stream_b) CLBM_CUDA(cudaStreamSynchronize(c->stream_b));
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    if (c->peer_err && *c->peer_err) {
        set_error("peer-memory ring: the wait for halo phase %d saw no signal from a neighbour within the time-out", *c->peer_err - 1);
        return CLBM_ESTATE;
    }
    return CLBM_OK;
}

int clbm_step_timed(clbm_ctx *c, int nsteps, float *ms)
{
    if (!c || !ms) { set_error("bad argument to clbm_step_timed"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    CLBM_CUDA(cudaEventRecord(c->ev0, c->stream));
    int rc = clbm_step(c, nsteps);
    if (rc) return rc;
    CLBM_CUDA(cudaEventRecord(c->ev1, c->stream));
    CLBM_CUDA(cudaEventSynchronize(c->ev1));
    CLBM_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return CLBM_OK;
}

int64_t clbm_launch_count(const clbm_ctx *c) { return c ? c->launches : -1; }

int clbm_profile_step(clbm_ctx *c, const char **names, float *ms, int cap)
{
    if (!c || !names || !ms || cap <= 0) { set_error("bad argument to clbm_profile_step"); return CLBM_EINVAL; }
    // an x-slab context can be profiled on a ring inside ONE process (every launch is followed by a host synchronisation, which
    // a ring across processes would not survive): the self ring of tools/self_ring_bench.py
    if (c->multi && !c->peer_mode) { set_error("profile a single-slab context or a context on a same-process peer ring"); return CLBM_ESTATE; }
    CLBM_CUDA(cudaSetDevice(c->device));
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    c->prof.clear();
    c->profiling = true;
    int rc = c->multi ? slab_step_for_profile(c) : model_step(c);
    c->profiling = false;
    if (rc) return rc;
    if (!c->multi) count_step(c);
    int n = 0;
    for (auto &k : c->prof) {
        if (n >= cap) break;
        names[n] = k.name.c_str();   // static strings behind std::string copies owned by ctx->prof
        ms[n] = k.ms;
        ++n;
    }
    return n;
}

int clbm_kernel_timing_begin(clbm_ctx *c, int cap)
{
    if (!c || cap < 1) { set_error("bad argument to clbm_kernel_timing_begin"); return CLBM_EINVAL; }
    for (auto e : c->kev) cudaEventDestroy(e);
    c->kev.clear();
    c->ktiming = true;
    c->ktiming_cap = cap;
    return CLBM_OK;
}

int clbm_kernel_timing_end(clbm_ctx *c, float *avg_ms, int *count, const char **kernel_name)
{
    if (!c || !avg_ms || !count) { set_error("bad argument to clbm_kernel_timing_end"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    c->ktiming = false;
    double sum = 0.0;
    int n = 0;
    for (size_t i = 0; i + 1 < c->kev.size(); i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->kev[i], c->kev[i + 1]) == cudaSuccess) { sum += ms; ++n; }
    }
    for (auto e : c->kev) cudaEventDestroy(e);
    c->kev.clear();
    *avg_ms = n ? (float)(sum / n) : 0.f;
    *count = n;
    if (kernel_name) *kernel_name = c->kname;
    return CLBM_OK;
}

int clbm_alloc_host(size_t bytes, void **ptr)
{
    if (!ptr) { set_error("null argument"); return CLBM_EINVAL; }
    *ptr = nullptr;
    // pinned pages are placed where the allocating thread runs: allocate from a CPU of the current GPU's NUMA node, so that
    // uploads do not cross the socket interconnect (best effort: a box without NUMA information allocates as before)
    NumaPin pin;
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
    pin.restore();
    if (e != cudaSuccess) return cuda_fail(e, "cudaHostAlloc", __FILE__, __LINE__);
    return CLBM_OK;
}
int clbm_free_host(void *ptr)
{
    if (ptr) CLBM_CUDA(cudaFreeHost(ptr));
    return CLBM_OK;
}

int clbm_reduce(clbm_ctx *c, int kind, double *out)
{
    if (!c || !out) { set_error("bad argument to clbm_reduce"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    return model_reduce(c, kind, out);
}

int clbm_diag_contact_angle(clbm_ctx *c, double rho_cut, int *base_y, int *base, int *height)
{
    if (!c) { set_error("bad argument to clbm_diag_contact_angle"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    return diag_contact_angle(c, rho_cut, base_y, base, height);
}

int clbm_diag_contact_angle_slab(clbm_ctx *c, double rho_cut, int base_y_in, int *out4)
{
    if (!c || !out4) { set_error("bad argument to clbm_diag_contact_angle_slab"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    return diag_contact_angle_raw(c, rho_cut, base_y_in, out4);
}

int clbm_diag_interface_heights(clbm_ctx *c, double phi_mid, int *y_at_x0, int *y_at_xmid)
{
    if (!c) { set_error("bad argument to clbm_diag_interface_heights"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    return diag_interface_heights(c, phi_mid, y_at_x0, y_at_xmid);
}

int clbm_halo_buffer(clbm_ctx *c, int phase, int side, int recv, void **dev_ptr, size_t *bytes)
{
    if (!c || phase < 0 || phase > 2 || side < 0 || side > 1 || recv < 0 || recv > 1 || !dev_ptr || !bytes) { set_error("bad argument to clbm_halo_buffer"); return CLBM_EINVAL; }
    if (!c->multi) { set_error("no halo buffers on a single-slab context"); return CLBM_ESTATE; }
    *dev_ptr = c->halo[phase][side][recv];
    *bytes = c->halo_bytes[phase];
    return CLBM_OK;
}

int clbm_halo_pack(clbm_ctx *c, int phase)
{
    if (!c || !c->multi) { set_error("halo pack needs an x-slab context"); return CLBM_ESTATE; }
    CLBM_CUDA(cudaSetDevice(c->device));
    return halo_pack(c, phase);
}
int clbm_halo_unpack(clbm_ctx *c, int phase)
{
    if (!c || !c->multi) { set_error("halo unpack needs an x-slab context"); return CLBM_ESTATE; }
    CLBM_CUDA(cudaSetDevice(c->device));
    return halo_unpack(c, phase);
}

void *clbm_stream(clbm_ctx *c) { return c ? (void *)c->stream : nullptr; }
void *clbm_boundary_stream(clbm_ctx *c)
{
    if (!c || !overlap_supported(c)) return nullptr;
    cudaSetDevice(c->device);
    if (ensure_boundary_stream(c)) return nullptr;
    return (void *)c->stream_b;
}

}  // extern "C"
