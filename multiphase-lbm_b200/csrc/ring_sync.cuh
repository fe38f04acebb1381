// ring_sync.cuh -- the flag protocol of the peer-memory ring (slab_comm.cu), as separate one-thread kernels or FUSED into the
// pack / unpack kernels of a halo phase.
//
// Every context owns a page of flag words in its mailbox.  A phase of the exchange is complete for a receiver when both of its
// neighbours have stored their sequence number of that phase into its arrive[] words -- which they do after everything they
// packed into its mailbox is visible system-wide.  Fused form: the LAST block of a pack kernel to finish bumps the sequence
// number and stores it into both neighbours' flags (every block fenced its stores before it took its ticket); every block of
// an unpack kernel polls the own arrive[] words before its first load, and the last block to finish bumps the wait counter.
// That takes the four one-thread launches of a step (signal, wait, signal, wait) out of the stream -- and measured SLOWER on the
// B200 (64-plane Shan-Chen slab: 1383 us per step against 1225 with the separate kernels; system-scope fences in every thread of
// the pack kernels and hundreds of polling blocks cost more than four tiny launches), so the fused form is an opt-in experiment
// (CLBM_RING_FUSE=1) and the separate kernels are the default.
#pragma once
#include <cuda_runtime.h>

namespace clbm {

struct MailFlags {
    unsigned arrive[3][2];   // [phase][side]: sequence number of the last block the neighbour on `side` completed in our mailbox
    unsigned seq[3];         // signals this context has sent, per phase
    unsigned expect[3];      // waits this context has done, per phase
    unsigned ticket[3][2];   // [phase][pack, unpack]: blocks of the running fused kernel that are done
};

struct RingSync {
    MailFlags *mine, *left, *right;
    int mode;                // 0: no synchronisation in this kernel, 1: signal when the grid is done (pack), 2: wait first (unpack)
    int phase;
    unsigned nblocks;        // blocks of the grid
    unsigned long long timeout_ns;
    int *err;                // mapped host word: phase + 1 when a wait gave up
};

__device__ __forceinline__ unsigned long long ring_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// spin until both neighbours have completed sequence number v of `phase` in our mailbox (one thread)
__device__ __forceinline__ void ring_spin(MailFlags *mine, int phase, unsigned v, unsigned long long timeout_ns, int *err)
{
    const unsigned long long t0 = ring_timer_ns();
    volatile unsigned *a0 = &mine->arrive[phase][0], *a1 = &mine->arrive[phase][1];
    unsigned spins = 0;
    while ((int)(*a0 - v) < 0 || (int)(*a1 - v) < 0) {
        if ((++spins & 1023u) == 0 && ring_timer_ns() - t0 > timeout_ns) {   // a neighbour died or never joined: report, do not hang the GPU
            *err = phase + 1;
            break;
        }
        __nanosleep(64);
    }
    __threadfence_system();
}

__device__ __forceinline__ void ring_send(MailFlags *mine, MailFlags *left, MailFlags *right, int phase)
{
    const unsigned v = mine->seq[phase] + 1u;
    mine->seq[phase] = v;
    __threadfence_system();   // everything packed into the neighbours' mailboxes is ordered before the flags
    *(volatile unsigned *)&left->arrive[phase][1] = v;    // we are the side-1 neighbour of our left neighbour
    *(volatile unsigned *)&right->arrive[phase][0] = v;
}

// first statement of a fused unpack kernel (all threads of every block)
__device__ __forceinline__ void ring_kernel_begin(const RingSync &r)
{
    if (r.mode != 2) return;
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
        // expect[] moves only when the whole grid is done, i.e. after every block has read it here
        const unsigned v = *(volatile unsigned *)&r.mine->expect[r.phase] + 1u;
        ring_spin(r.mine, r.phase, v, r.timeout_ns, r.err);
    }
    __syncthreads();
}

// last statement of a fused pack / unpack kernel (all threads of every block, no early return before it)
__device__ __forceinline__ void ring_kernel_end(const RingSync &r)
{
    if (r.mode == 0) return;
    // this thread's stores (possibly into a neighbour's mailbox) before the block's ticket: a device-scope fence per thread; the
    // last block adds the system-scope fence in ring_send, which is cumulative over everything the tickets made visible to it
    // (a system-scope fence in every thread of a 10 000-block pack kernel is what made the first fused form slow)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
        unsigned *t = &r.mine->ticket[r.phase][r.mode - 1];
        if (atomicAdd(t, 1u) == r.nblocks - 1u) {   // the last block of the grid
            *t = 0u;
            __threadfence();
            if (r.mode == 1) ring_send(r.mine, r.left, r.right, r.phase);
            else *(volatile unsigned *)&r.mine->expect[r.phase] = r.mine->expect[r.phase] + 1u;
        }
    }
}

}  // namespace clbm
