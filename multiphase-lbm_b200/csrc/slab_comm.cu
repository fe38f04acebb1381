// slab_comm.cu -- the x-slab ring driven entirely from the library: clbm_slab_step(ctx, n) runs n slab steps (stages +
// both ghost exchanges) without returning to the caller.
//
// Two transports, chosen per context:
//
//  * PEER MEMORY (default on one node; clbm_peer_export / clbm_peer_connect, or clbm_peer_connect_local for contexts of one
//    process).  All halo blocks of a context live in one allocation, the mailbox (support_kernels.cu: halo_alloc), which is
//    exported as ONE cudaIpcMemHandle.  After the connect, the pack step of a phase writes straight into the receive block
//    of the neighbour (plane copies for the moment halo, the pack kernel's stores for the crossing populations: the data
//    crosses NVLink exactly once, no staging copy, no SM-resident library kernel).  An exchange is then two one-thread
//    kernels: SIGNAL bumps this rank's sequence number of the phase and stores it into both neighbours' flag words (after a
//    system-scope fence, behind the pack in stream order); WAIT spins until both of our own flag words have reached the
//    number of waits done so far.  The counters live in device memory, so the same kernels replay inside a CUDA graph.
//    Why a receive block may be overwritten: the neighbour consumed phase p of step n before it signalled the NEXT phase,
//    and we wait for that signal before we pack phase p of step n+1 (sequence: unpack p -> ... -> signal p' on its stream).
//
//  * NCCL send/recv groups (clbm_comm_unique_id / clbm_comm_init): the fallback when the mailboxes cannot be mapped.  NCCL is
//    resolved at run time (dlopen("libnccl.so.2")); the few types used are declared here, so libclbm.so needs neither NCCL
//    headers nor the library to build or to load.
//
// On a peer ring two consecutive steps (after which every buffer pointer is back where it started) are captured ONCE per
// starting parity into a CUDA graph -- both streams of the overlap protocol, the plane copies, the signal / wait kernels --
// and replayed n/2 times: BASELINE configs[2] on 8 GPUs is a 0.17 ms kernel per step, which a dozen separate launches and
// copies per step cannot feed from the host.
#include <dlfcn.h>
#include <unistd.h>

#include <cstring>
#include <mutex>

#include "clbm_internal.h"
#include "ring_sync.cuh"

namespace clbm {

// ---- NCCL, resolved at run time ---------------------------------------------------------------------------------
// (declared locally: the ABI of these eight entry points and of the two enums has been stable since NCCL 2.7)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                 // ncclSuccess == 0
static const int NCCL_UINT8 = 1;          // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
    bool ok;
    char why[200];
};

static NcclApi *nccl_api()
{
    static NcclApi api = {};
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { snprintf(api.why, sizeof(api.why), "NCCL not found (dlopen libnccl.so.2): %s", dlerror()); return; }
        bool all = true;
        auto sym = [&](const char *name) { void *p = dlsym(h, name); if (!p) all = false; return p; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        if (!all) { snprintf(api.why, sizeof(api.why), "libnccl.so.2 lacks a point-to-point symbol (NCCL >= 2.7 needed)"); return; }
        api.ok = true;
    });
    if (!api.ok) { set_error("%s", api.why); return nullptr; }   // every failed caller gets the reason, not only the first
    return &api;
}

#define CLBM_NCCL(api, call)                                                                          \
    do {                                                                                              \
        ncclResult_t r__ = (call);                                                                    \
        if (r__ != 0) { set_error("NCCL: %s (%s)", (api)->GetErrorString(r__), #call); return CLBM_ECUDA; } \
    } while (0)

// one exchange phase with both ring neighbours on stream `st`: my side-0 send buffer travels to the left neighbour's side-1
// receive buffer and vice versa (the pairing of slab.ring_exchange; with two ranks left == right and NCCL matches the two
// send/recv pairs of a group in posting order, which is the same on both ranks)
static int nccl_exchange(clbm_ctx *c, int phase, cudaStream_t st)
{
    const size_t n = c->halo_bytes[phase];
    void *send0 = c->halo[phase][0][0], *send1 = c->halo[phase][1][0], *recv0 = c->halo[phase][0][1], *recv1 = c->halo[phase][1][1];
    const int R = c->comm_size, r = c->comm_rank;
    NcclApi *a = nccl_api();
    if (!a) return CLBM_ESTATE;
    ncclComm_t comm = (ncclComm_t)c->comm;
    const int left = (r - 1 + R) % R, right = (r + 1) % R;
    CLBM_NCCL(a, a->GroupStart());
    // a failure between GroupStart and GroupEnd must still close the group: an open group swallows every later NCCL call
    // of the process (torch's included)
    ncclResult_t bad = 0;
    const char *what = "";
    auto step = [&](ncclResult_t r_, const char *w) { if (r_ != 0 && bad == 0) { bad = r_; what = w; } };
    step(a->Send(send0, n, NCCL_UINT8, left, comm, st), "ncclSend(left)");
    step(a->Recv(recv1, n, NCCL_UINT8, right, comm, st), "ncclRecv(right)");
    step(a->Send(send1, n, NCCL_UINT8, right, comm, st), "ncclSend(right)");
    step(a->Recv(recv0, n, NCCL_UINT8, left, comm, st), "ncclRecv(left)");
    const ncclResult_t end = a->GroupEnd();
    if (bad != 0) { set_error("NCCL: %s (%s)", a->GetErrorString(bad), what); return CLBM_ECUDA; }
    if (end != 0) { set_error("NCCL: %s (ncclGroupEnd)", a->GetErrorString(end)); return CLBM_ECUDA; }
    return 0;
}

// ---- peer-memory ring -------------------------------------------------------------------------------------------
static MailFlags *flags_of(void *mailbox, const clbm_ctx *c) { return (MailFlags *)((char *)mailbox + c->mailbox_flags_off); }

__global__ void peer_signal_kernel(MailFlags *mine, MailFlags *left, MailFlags *right, int phase) { ring_send(mine, left, right, phase); }

__global__ void peer_wait_kernel(MailFlags *mine, int phase, unsigned long long timeout_ns, int *err)
{
    const unsigned v = mine->expect[phase] + 1u;
    mine->expect[phase] = v;
    ring_spin(mine, phase, v, timeout_ns, err);
}

unsigned long long peer_timeout_ns()
{
    static const int ms = env_int("CLBM_PEER_TIMEOUT_MS", 20000);
    return (unsigned long long)(ms > 0 ? ms : 20000) * 1000000ull;
}

// what a fused pack (mode 1) / unpack (mode 2) kernel of `phase` needs; mode 0 when the context's ring is not fused right now
RingSync ring_sync_for(const clbm_ctx *c, int phase, int mode, unsigned nblocks)
{
    RingSync r = {};
    if (!c->ring_fuse || !c->peer_mode) return r;
    if (c->ring_fuse == 2 && mode == 2) return r;   // signals fused, waits as separate kernels
    r.mine = flags_of(c->mailbox, c);
    r.left = flags_of(c->peer_base[0], c);
    r.right = flags_of(c->peer_base[1], c);
    r.mode = mode;
    r.phase = phase;
    r.nblocks = nblocks;
    r.timeout_ns = peer_timeout_ns();
    r.err = c->peer_err;
    return r;
}

static int peer_exchange(clbm_ctx *c, int phase, cudaStream_t st)
{
    MailFlags *mine = flags_of(c->mailbox, c);
    {
        LaunchScope ls(c, "peer_signal");
        peer_signal_kernel<<<1, 1, 0, st>>>(mine, flags_of(c->peer_base[0], c), flags_of(c->peer_base[1], c), phase);
        CLBM_CUDA(cudaGetLastError());
    }
    {
        LaunchScope ls(c, "peer_wait");
        peer_wait_kernel<<<1, 1, 0, st>>>(mine, phase, peer_timeout_ns(), c->peer_err);
        CLBM_CUDA(cudaGetLastError());
    }
    return 0;
}

static int ring_exchange(clbm_ctx *c, int phase, cudaStream_t st)
{
    if (c->peer_mode) return peer_exchange(c, phase, st);
    if (c->comm) return nccl_exchange(c, phase, st);
    set_error("this context has no ring (clbm_peer_connect or clbm_comm_init first)");
    return CLBM_ESTATE;
}

struct PeerHandle {                // what clbm_peer_export writes: CLBM_PEER_HANDLE_BYTES
    cudaIpcMemHandle_t mem;        // 64 bytes
    unsigned long long magic, bytes, flags_off, psi_off;
    int pid, device, nx, model;
    char pad[CLBM_PEER_HANDLE_BYTES - 64 - 4 * 8 - 4 * 4];
};
static_assert(sizeof(PeerHandle) == CLBM_PEER_HANDLE_BYTES, "handle size");
static const unsigned long long PEER_MAGIC = 0x434c424d50454552ull;   // "CLBMPEER"

static int peer_common_init(clbm_ctx *c)
{
    if (!c->peer_err) {
        CLBM_CUDA(cudaHostAlloc((void **)&c->peer_err, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
        *c->peer_err = 0;
    }
    // the flag page restarts from zero on both ends of every link (each rank zeroes its own; the caller's barrier between
    // connect and the first step orders that before any neighbour's signal)
    CLBM_CUDA(cudaMemsetAsync((char *)c->mailbox + c->mailbox_flags_off, 0, 4096, c->stream));
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

static void drop_graphs(clbm_ctx *c)
{
    for (auto &g : c->slab_graph)
        if (g) { cudaGraphExecDestroy((cudaGraphExec_t)g); g = nullptr; }
    c->slab_graph_failed = 0;
}

// one slab step issued call by call (both protocols); also what the graph capture records
static int slab_step_eager(clbm_ctx *c)
{
    const bool overlap = clbm_overlap_supported(c) != 0;
    cudaStream_t xs = c->stream;
    if (overlap) {
        xs = (cudaStream_t)clbm_boundary_stream(c);
        if (!xs) return CLBM_ECUDA;
    }
    const int s0 = overlap ? 10 : 0;
    int rc;
    if (c->peer_mode && c->env.ring_fuse != 0 && c->env.ring_fuse != 1) {
        // default on a peer ring: the SIGNAL of a phase rides on the last block of its pack kernel (two launches less per step);
        // the waits stay one-thread kernels (hundreds of polling blocks in the unpack kernels measured slower)
        cudaStream_t w0 = clbm_overlap_variant(c) == 1 ? xs : c->stream;
        auto wait = [&](int phase, cudaStream_t st) -> int {
            LaunchScope ls(c, "peer_wait");
            peer_wait_kernel<<<1, 1, 0, st>>>(flags_of(c->mailbox, c), phase, peer_timeout_ns(), c->peer_err);
            CLBM_CUDA(cudaGetLastError());
            return 0;
        };
        c->ring_fuse = 2;
        rc = clbm_step_stage(c, s0);
        if (!rc) rc = wait(0, w0);
        if (!rc) rc = clbm_step_stage(c, s0 + 1);
        if (!rc) rc = wait(1, xs);
        if (!rc) rc = clbm_step_stage(c, s0 + 2);
        c->ring_fuse = 0;
        return rc;
    }
    if (c->peer_mode && c->env.ring_fuse == 1) {
        // CLBM_RING_FUSE=1 (an experiment that stays off: measured SLOWER): the signal rides on the last
        // block of every pack kernel, the wait on the first instruction of every unpack kernel
        c->ring_fuse = 1;
        rc = clbm_step_stage(c, s0);
        if (!rc) rc = clbm_step_stage(c, s0 + 1);
        if (!rc) rc = clbm_step_stage(c, s0 + 2);
        c->ring_fuse = 0;
        return rc;
    }
    if ((rc = clbm_step_stage(c, s0))) return rc;
    if ((rc = ring_exchange(c, 0, clbm_overlap_variant(c) == 1 ? xs : c->stream))) return rc;   // form 1 moves the moment halo on the boundary stream
    if ((rc = clbm_step_stage(c, s0 + 1))) return rc;
    if ((rc = ring_exchange(c, 1, xs))) return rc;
    return clbm_step_stage(c, s0 + 2);
}

// capture the next two steps (starting at the current parity) into a graph; the host-side state (parity, step count)
// advances during the capture exactly as if the steps had run, so the caller launches the graph once for them
static int capture_two_steps(clbm_ctx *c, cudaGraphExec_t *out, int64_t *launches_per_replay)
{
    const int64_t l0 = c->launches;
    if (clbm_overlap_supported(c) && !clbm_boundary_stream(c)) return CLBM_ECUDA;   // streams / events exist before the capture
    CLBM_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
    int rc = slab_step_eager(c);
    if (!rc) rc = slab_step_eager(c);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess || !graph) return cuda_fail(e, "cudaStreamEndCapture (slab step)", __FILE__, __LINE__);
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate (slab step)", __FILE__, __LINE__);
    *out = exec;
    *launches_per_replay = c->launches - l0;
    return 0;
}

// one call-by-call slab step for clbm_profile_step (every launch bracketed by events, host-synchronised: a ring of one process)
int slab_step_for_profile(clbm_ctx *c) { return slab_step_eager(c); }

}  // namespace clbm

using namespace clbm;

extern "C" {

int clbm_comm_unique_id(void *id128)
{
    if (!id128) { set_error("null argument"); return CLBM_EINVAL; }
    NcclApi *a = nccl_api();
    if (!a) return CLBM_ESTATE;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    CLBM_NCCL(a, a->GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return CLBM_OK;
}

int clbm_comm_init(clbm_ctx *c, const void *id128, int rank, int nranks)
{
    if (!c || !id128 || nranks < 2 || rank < 0 || rank >= nranks) { set_error("bad argument to clbm_comm_init (a ring has at least two ranks)"); return CLBM_EINVAL; }
    if (!c->multi) { set_error("clbm_comm_init needs an x-slab context (nx < nx_global)"); return CLBM_ESTATE; }
    if (c->comm) { set_error("this context already has a communicator"); return CLBM_ESTATE; }
    if (c->peer_mode) { set_error("this context is on a peer-memory ring"); return CLBM_ESTATE; }
    NcclApi *a = nccl_api();
    if (!a) return CLBM_ESTATE;
    CLBM_CUDA(cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    CLBM_NCCL(a, a->CommInitRank(&comm, nranks, id, rank));
    c->comm = comm;
    c->comm_rank = rank;
    c->comm_size = nranks;
    return CLBM_OK;
}

int clbm_comm_destroy(clbm_ctx *c)
{
    if (!c) return CLBM_OK;
    if (c->comm) {
        NcclApi *a = nccl_api();
        if (a) {
            cudaSetDevice(c->device);
            cudaStreamSynchronize(c->stream);
            if (c->stream_b) cudaStreamSynchronize(c->stream_b);
            a->CommDestroy((ncclComm_t)c->comm);
        }
        c->comm = nullptr;
    }
    c->comm_size = 0;
    return CLBM_OK;
}

int clbm_peer_export(clbm_ctx *c, void *handle)
{
    if (!c || !handle) { set_error("bad argument to clbm_peer_export"); return CLBM_EINVAL; }
    if (!c->multi || !c->mailbox) { set_error("clbm_peer_export needs an x-slab context (nx < nx_global)"); return CLBM_ESTATE; }
    CLBM_CUDA(cudaSetDevice(c->device));
    PeerHandle h;
    memset(&h, 0, sizeof(h));
    CLBM_CUDA(cudaIpcGetMemHandle(&h.mem, c->mailbox));
    h.magic = PEER_MAGIC;
    h.bytes = c->mailbox_bytes;
    h.flags_off = c->mailbox_flags_off;
    h.psi_off = c->mailbox_psi_off;
    h.pid = (int)getpid();
    h.device = c->device;
    h.nx = c->geo.nx;
    h.model = c->prm.model;
    memcpy(handle, &h, sizeof(h));
    return CLBM_OK;
}

int clbm_peer_connect(clbm_ctx *c, const void *left_handle, const void *right_handle)
{
    if (!c || !left_handle || !right_handle) { set_error("bad argument to clbm_peer_connect"); return CLBM_EINVAL; }
    if (!c->multi || !c->mailbox) { set_error("clbm_peer_connect needs an x-slab context (nx < nx_global)"); return CLBM_ESTATE; }
    if (c->peer_mode || c->comm) { set_error("this context already has a ring"); return CLBM_ESTATE; }
    CLBM_CUDA(cudaSetDevice(c->device));
    PeerHandle h[2];
    memcpy(&h[0], left_handle, sizeof(PeerHandle));
    memcpy(&h[1], right_handle, sizeof(PeerHandle));
    for (int s = 0; s < 2; ++s) {
        if (h[s].magic != PEER_MAGIC) { set_error("not a clbm_peer_export handle"); return CLBM_EINVAL; }
        if (h[s].flags_off != c->mailbox_flags_off || h[s].psi_off != c->mailbox_psi_off || h[s].model != c->prm.model) {
            set_error("neighbour mailbox layout differs (flag page at %llu / %llu): ring members must share the model and the y, z extents", h[s].flags_off, (unsigned long long)c->mailbox_flags_off);
            return CLBM_EINVAL;
        }
        if (h[s].pid == (int)getpid()) { set_error("neighbour lives in this process: use clbm_peer_connect_local"); return CLBM_EINVAL; }
    }
    void *base[2] = {nullptr, nullptr};
    cudaError_t e = cudaIpcOpenMemHandle(&base[0], h[0].mem, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaIpcOpenMemHandle (left neighbour, device %d): %s", h[0].device, cudaGetErrorString(e)); return CLBM_ECUDA; }
    if (memcmp(&h[0].mem, &h[1].mem, sizeof(cudaIpcMemHandle_t)) == 0) base[1] = base[0];   // ring of two: one neighbour, mapped once
    else {
        e = cudaIpcOpenMemHandle(&base[1], h[1].mem, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); cudaIpcCloseMemHandle(base[0]); set_error("cudaIpcOpenMemHandle (right neighbour, device %d): %s", h[1].device, cudaGetErrorString(e)); return CLBM_ECUDA; }
    }
    int rc = peer_common_init(c);
    if (rc) { cudaIpcCloseMemHandle(base[0]); if (base[1] != base[0]) cudaIpcCloseMemHandle(base[1]); return rc; }
    c->peer_base[0] = base[0];
    c->peer_base[1] = base[1];
    c->peer_nx[0] = h[0].nx;
    c->peer_nx[1] = h[1].nx;
    c->halo0_direct = c->fld0_in_mailbox;
    c->peer_mode = 1;
    drop_graphs(c);
    return CLBM_OK;
}

int clbm_peer_connect_local(clbm_ctx *c, clbm_ctx *left, clbm_ctx *right)
{
    if (!c || !left || !right) { set_error("bad argument to clbm_peer_connect_local"); return CLBM_EINVAL; }
    if (!c->multi || !c->mailbox || !left->mailbox || !right->mailbox) { set_error("clbm_peer_connect_local needs x-slab contexts"); return CLBM_ESTATE; }
    if (c->peer_mode || c->comm) { set_error("this context already has a ring"); return CLBM_ESTATE; }
    clbm_ctx *nb[2] = {left, right};
    CLBM_CUDA(cudaSetDevice(c->device));
    for (int s = 0; s < 2; ++s) {
        if (nb[s]->mailbox_flags_off != c->mailbox_flags_off || nb[s]->mailbox_psi_off != c->mailbox_psi_off || nb[s]->prm.model != c->prm.model) { set_error("neighbour mailbox layout differs"); return CLBM_EINVAL; }
        if (nb[s]->device != c->device) {
            int can = 0;
            CLBM_CUDA(cudaDeviceCanAccessPeer(&can, c->device, nb[s]->device));
            if (!can) { set_error("device %d cannot access device %d", c->device, nb[s]->device); return CLBM_ESTATE; }
            cudaError_t e = cudaDeviceEnablePeerAccess(nb[s]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
            cudaGetLastError();
        }
    }
    int rc = peer_common_init(c);
    if (rc) return rc;
    c->peer_base[0] = left->mailbox;
    c->peer_base[1] = right->mailbox;
    c->peer_nx[0] = left->geo.nx;
    c->peer_nx[1] = right->geo.nx;
    c->halo0_direct = c->fld0_in_mailbox;
    c->peer_mode = 2;
    drop_graphs(c);
    return CLBM_OK;
}

int clbm_peer_disconnect(clbm_ctx *c)
{
    if (!c) return CLBM_OK;
    if (c->peer_mode) {
        cudaSetDevice(c->device);
        if (c->stream) cudaStreamSynchronize(c->stream);
        if (c->stream_b) cudaStreamSynchronize(c->stream_b);
        drop_graphs(c);
        if (c->peer_mode == 1) {
            if (c->peer_base[0]) cudaIpcCloseMemHandle(c->peer_base[0]);
            if (c->peer_base[1] && c->peer_base[1] != c->peer_base[0]) cudaIpcCloseMemHandle(c->peer_base[1]);
        }
        c->peer_base[0] = c->peer_base[1] = nullptr;
        c->peer_mode = 0;
        c->halo0_direct = 0;
    }
    if (c->peer_err) { cudaFreeHost(c->peer_err); c->peer_err = nullptr; }
    return CLBM_OK;
}

// the two halves of an exchange on a peer ring, for callers that drive the stages themselves: SIGNAL after the pack of the
// phase, WAIT before its unpack, both on the stream the protocol uses (boundary != 0: the boundary stream of the overlap protocol).
// Several contexts of ONE process sharing a GPU must issue every signal of a phase before any wait of that phase: a spinning
// wait kernel may sit in front of another context's signal in a hardware queue the two streams happen to share.
int clbm_slab_signal(clbm_ctx *c, int phase, int boundary)
{
    if (!c || phase < 0 || phase > 2) { set_error("bad argument to clbm_slab_signal"); return CLBM_EINVAL; }
    if (!c->multi || !c->peer_mode) { set_error("clbm_slab_signal needs an x-slab context on a peer-memory ring"); return CLBM_ESTATE; }
    CLBM_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = boundary ? (cudaStream_t)clbm_boundary_stream(c) : c->stream;
    if (!st) { set_error("no boundary stream on this context"); return CLBM_ESTATE; }
    LaunchScope ls(c, "peer_signal");
    peer_signal_kernel<<<1, 1, 0, st>>>(flags_of(c->mailbox, c), flags_of(c->peer_base[0], c), flags_of(c->peer_base[1], c), phase);
    CLBM_CUDA(cudaGetLastError());
    return CLBM_OK;
}

int clbm_slab_wait(clbm_ctx *c, int phase, int boundary)
{
    if (!c || phase < 0 || phase > 2) { set_error("bad argument to clbm_slab_wait"); return CLBM_EINVAL; }
    if (!c->multi || !c->peer_mode) { set_error("clbm_slab_wait needs an x-slab context on a peer-memory ring"); return CLBM_ESTATE; }
    CLBM_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = boundary ? (cudaStream_t)clbm_boundary_stream(c) : c->stream;
    if (!st) { set_error("no boundary stream on this context"); return CLBM_ESTATE; }
    LaunchScope ls(c, "peer_wait");
    peer_wait_kernel<<<1, 1, 0, st>>>(flags_of(c->mailbox, c), phase, peer_timeout_ns(), c->peer_err);
    CLBM_CUDA(cudaGetLastError());
    return CLBM_OK;
}

int clbm_ring_kind(const clbm_ctx *c) { return !c ? 0 : (c->peer_mode ? 2 : (c->comm ? 1 : 0)); }

// pack + exchange + unpack of one halo phase on the launching stream (phase 2: the node mask after an upload; phase 0 after
// the moments: the ghost values a field download needs).  Every rank of the ring calls it; asynchronous.
int clbm_slab_exchange(clbm_ctx *c, int phase)
{
    if (!c || phase < 0 || phase > 2) { set_error("bad argument to clbm_slab_exchange"); return CLBM_EINVAL; }
    if (!c->multi || !(c->peer_mode || c->comm)) { set_error("clbm_slab_exchange needs an x-slab context on a ring"); return CLBM_ESTATE; }
    CLBM_CUDA(cudaSetDevice(c->device));
    if (c->stream_b) {   // whatever the boundary stream still has in flight comes first (device-side order, no host sync)
        CLBM_CUDA(cudaEventRecord(c->ev_b, c->stream_b));
        CLBM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_b, 0));
    }
    int rc;
    if ((rc = clbm_halo_pack(c, phase))) return rc;
    if ((rc = ring_exchange(c, phase, c->stream))) return rc;
    return clbm_halo_unpack(c, phase);
}

// nsteps slab steps; asynchronous like clbm_step (any download / reduce / clbm_sync synchronises).  Uses the boundary-first
// overlap protocol (stages 10-12, exchanges on the boundary stream) where the context supports it, else the sequential one
// (stages 0-2, exchanges on the launching stream).  Every rank of the ring must call it with the same nsteps.
int clbm_slab_step(clbm_ctx *c, int nsteps)
{
    if (!c || nsteps < 0) { set_error("bad argument to clbm_slab_step"); return CLBM_EINVAL; }
    if (!c->multi || !(c->peer_mode || c->comm)) { set_error("clbm_slab_step needs an x-slab context and clbm_peer_connect / clbm_comm_init first"); return CLBM_ESTATE; }
    CLBM_CUDA(cudaSetDevice(c->device));
    int rc;
    int left = nsteps;
    // the first two steps of a context always run call by call: they set the per-device kernel attributes and encode the
    // tensor maps of both parities, none of which belongs inside a capture
    const bool graphs = c->peer_mode && c->env.slab_graph != 0 && !c->ktiming && !c->profiling && !c->slab_graph_failed;
    while (left > 0) {
        if (graphs && left >= 2 && c->steps_taken >= 2) {
            const int par = c->parity;
            if (!c->slab_graph[par]) {
                cudaGraphExec_t exec = nullptr;
                int64_t per = 0;
                const int prc = capture_two_steps(c, &exec, &per);
                if (prc) {
                    // nothing was launched, but parity / step count may have moved: a failed capture is not recoverable here
                    c->slab_graph_failed = 1;
                    return prc;
                }
                c->slab_graph[par] = exec;
                c->slab_graph_launches[par] = per;
            } else {
                c->steps_taken += 2;
                c->launches += c->slab_graph_launches[par];
            }
            CLBM_CUDA(cudaGraphLaunch((cudaGraphExec_t)c->slab_graph[par], c->stream));
            left -= 2;
            continue;
        }
        if ((rc = slab_step_eager(c))) return rc;
        --left;
    }
    return CLBM_OK;
}

}  // extern "C"
