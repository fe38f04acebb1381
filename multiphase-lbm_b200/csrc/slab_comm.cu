// slab_comm.cu -- the x-slab ring driven entirely from the library: clbm_slab_step(ctx, n) runs n slab steps (stages +
// both ghost exchanges) without returning to the caller, the exchanges being ncclSend / ncclRecv groups issued on the
// library's own streams.
//
// Why: BASELINE configs[2] at 8 GPUs is 256 columns per GPU, a 0.17 ms kernel.  Driven from Python (slab.DistRing: three
// clbm_step_stage calls and two torch batch_isend_irecv calls per step) the step is HOST-launch bound (0.23-0.27 ms,
// DESIGN.md section 4).  Here a step costs the host 7 kernel launches, 4 small copies and two NCCL groups, and the NCCL
// kernels run on the boundary stream itself (high priority), not on a process-group stream.
//
// NCCL is resolved at run time (dlopen("libnccl.so.2"): the copy torch already loaded, or the system one), so libclbm.so
// has no link-time dependency on it and single-GPU users never touch it.  One process per GPU; the communicator is built
// from a ncclUniqueId that the caller broadcasts (clbm_comm_unique_id on rank 0 -> any transport -> clbm_comm_init).
//
// STATUS: compiled and symbol-checked in the CPU suite; opt-in (slab.DistRing(native=True) / CLBM_SLAB_NATIVE=1) until it has
// been run on a multi-GPU box (tools/slab_check.py --native compares it bit-for-bit with the single slab).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "clbm_internal.h"

namespace clbm {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
    bool ok;
};

static NcclApi *nccl_api()
{
    static NcclApi api = {};
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("NCCL not found (dlopen libnccl.so.2): %s", dlerror()); return nullptr; }
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
    if (!all) { set_error("libnccl.so.2 lacks a point-to-point symbol (NCCL >= 2.7 needed)"); return nullptr; }
    api.ok = true;
    return &api;
}

#define CLBM_NCCL(api, call)                                                                          \
    do {                                                                                              \
        ncclResult_t r__ = (call);                                                                    \
        if (r__ != ncclSuccess) { set_error("NCCL: %s (%s)", (api)->GetErrorString(r__), #call); return CLBM_ECUDA; } \
    } while (0)

// one exchange phase with both ring neighbours on stream `st`: my side-0 send buffer travels to the left neighbour's side-1
// receive buffer and vice versa (the pairing of slab.ring_exchange; with two ranks left == right and NCCL matches the two
// send/recv pairs of a group in posting order, which is the same on both ranks)
static int ring_exchange(clbm_ctx *c, int phase, cudaStream_t st)
{
    const size_t n = c->halo_bytes[phase];
    void *send0 = c->halo[phase][0][0], *send1 = c->halo[phase][1][0], *recv0 = c->halo[phase][0][1], *recv1 = c->halo[phase][1][1];
    const int R = c->comm_size, r = c->comm_rank;
    NcclApi *a = nccl_api();
    if (!a) return CLBM_ESTATE;
    ncclComm_t comm = (ncclComm_t)c->comm;
    const int left = (r - 1 + R) % R, right = (r + 1) % R;
    CLBM_NCCL(a, a->GroupStart());
    CLBM_NCCL(a, a->Send(send0, n, ncclUint8, left, comm, st));
    CLBM_NCCL(a, a->Recv(recv1, n, ncclUint8, right, comm, st));
    CLBM_NCCL(a, a->Send(send1, n, ncclUint8, right, comm, st));
    CLBM_NCCL(a, a->Recv(recv0, n, ncclUint8, left, comm, st));
    CLBM_NCCL(a, a->GroupEnd());
    return 0;
}

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

// nsteps slab steps; asynchronous like clbm_step (any download / reduce / clbm_sync synchronises).  Uses the boundary-first
// overlap protocol (stages 10-12, exchanges on the boundary stream) where the context supports it, else the sequential one
// (stages 0-2, exchanges on the launching stream).  Every rank of the ring must call it with the same nsteps.
int clbm_slab_step(clbm_ctx *c, int nsteps)
{
    if (!c || nsteps < 0) { set_error("bad argument to clbm_slab_step"); return CLBM_EINVAL; }
    if (!c->multi || !c->comm) { set_error("clbm_slab_step needs an x-slab context and clbm_comm_init first"); return CLBM_ESTATE; }
    CLBM_CUDA(cudaSetDevice(c->device));
    const bool overlap = clbm_overlap_supported(c) != 0;
    cudaStream_t xs = c->stream;
    if (overlap) {
        xs = (cudaStream_t)clbm_boundary_stream(c);
        if (!xs) return CLBM_ECUDA;
    }
    const int s0 = overlap ? 10 : 0;
    int rc;
    for (int s = 0; s < nsteps; ++s) {
        if ((rc = clbm_step_stage(c, s0))) return rc;
        if ((rc = ring_exchange(c, 0, xs))) return rc;
        if ((rc = clbm_step_stage(c, s0 + 1))) return rc;
        if ((rc = ring_exchange(c, 1, xs))) return rc;
        if ((rc = clbm_step_stage(c, s0 + 2))) return rc;
    }
    return CLBM_OK;
}

}  // extern "C"
