// pulsatile.cu -- compliant-vessel case "Abbashub LBM/apps/PulsatileBloodFlow2D.h" (AB/ below) on the device:
// pressure-based D2Q9 MRT collision, Bouzidi curved moving walls, pull streaming, Zou/He pressure inlet/outlet,
// pressure-driven wall motion with fresh-node filling.  Everything the reference does per iteration
// (AB:764-790) -- including the parts it runs serially on the host -- is a kernel here, so a run never leaves HBM.
//
// This translation unit is compiled with -fmad=false: every +,-,*,/ is a correctly rounded IEEE operation in the
// reference's order, so the node mask (flag = Fobj >= 1, driven by the pressure through the wall ODE) and the
// fields stay BIT-IDENTICAL to the reference; the path is HBM-bound, the missing FMAs cost nothing.
//
// Device layout == reference layout: lattice[p*npop + k*nelem + i], i = y + ny*x, plus one zero pad double, because
// the reference's pull with wrapY == identity (AB:90) reads flat indices -1 / nelem that spill into
// the neighbouring population array; those reads are reproduced literally.
//
// Kernels (one time step, AB:764-790):
//   puls_fused     collide of all fluid nodes + pull stream + moments of the INTERIOR fluid nodes in one column-marching
//                  pass (default; CLBM_PULS_FUSED=0 selects the two-pass form puls_collide + puls_stream on all nodes)
//   puls_collide   MRT_Collision (AB:533-541) on fluid nodes: out = in - M^-1 S M (in - geq(P, Ux, Uy))
//   puls_bouzidi   border-node discovery (AB:294-382) + Bouzidi_quadratic (AB:553-601), one thread per column, one launch per wall;
//                  the Delta arrays are recomputed from the wall positions instead of being stored
//   puls_stream    Streaming (AB:603-616) + Inlet/Outlet_ZouHe (AB:618-669) + Macroscopic_Properties_g (AB:216-230)
//   puls_walls     Calculate_Pressure_and_Move_Walls (AB:243-272)
//   puls_fobj      Update_Fobj_for_Vessel_Walls + Fill_Fluid_Node + Fresh_Macroscopic_Values (AB:384-498), one thread
//                  per column; Fobj (AB:275-285) is a pure function of the wall positions and is never stored
//   puls_seed      Seed_From_Nearest_Fluid (AB:418-458) for fresh nodes of a stretch that opens up, in sweep order
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "clbm_internal.h"
#include "tma.cuh"

namespace clbm {
namespace puls {

struct Geo {
    int nx, ny, Y0;
    long long nelem, npop;
    double c;          // Y0 + 0.5
};
struct Par {
    double Rho0, S[9];
    double alpha, p_tissue;
};

// AB:29-38 k ordering; AB:41-49 "I" ordering (0 rest, 1 E, 2 N, 3 W, 4 S, 5 NE, 6 NW, 7 SW, 8 SE)
__host__ __device__ constexpr int ckx(int k) { constexpr int v[9] = {-1, 0, -1, -1, 0, 1, 0, 1, 1}; return v[k]; }
__host__ __device__ constexpr int cky(int k) { constexpr int v[9] = {0, -1, -1, 1, 0, 0, 1, 1, -1}; return v[k]; }
__host__ __device__ constexpr double tk(int k)
{
    constexpr double v[9] = {1. / 9., 1. / 9., 1. / 36., 1. / 36., 4. / 9., 1. / 9., 1. / 9., 1. / 36., 1. / 36.};
    return v[k];
}
__host__ __device__ constexpr int exI(int I) { constexpr int v[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}; return v[I]; }
__host__ __device__ constexpr int eyI(int I) { constexpr int v[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1}; return v[I]; }
__host__ __device__ constexpr int jbI(int I) { constexpr int v[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6}; return v[I]; }
__host__ __device__ constexpr int kfromI(int I) { constexpr int v[9] = {4, 5, 6, 0, 1, 7, 3, 2, 8}; return v[I]; }

// AB:501-507.  Rho0 is 1 in the reference (hsize = 1, AB:97-98): Rho0 / 3.0 is then the constant RN(1/3) and no division runs.
__host__ __device__ inline void equilibrium_g(double Rho0, double P, double U, double V, double geq[9])
{
    const double U2 = U * U + V * V;
    const double r3 = (Rho0 == 1.0) ? (1.0 / 3.0) : Rho0 / 3.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const double eU = ckx(k) * U + cky(k) * V;
        geq[k] = tk(k) * (P + r3 * (eU * (3.0 + 4.5 * eU) - 1.5 * U2));
    }
}

// x / C for a small integer constant C, CORRECTLY ROUNDED like the IEEE division it replaces, in three instructions
// (Markstein: q0 = RN(x r), rem = x - C q0 exactly by FMA, q = RN(q0 + rem r), r = RN(1/C)); an IEEE FP64 division is a
// ~30-instruction sequence with a branch, and reconvert() has ten of them per node.  Checked against x / C on 3e8
// operands per constant (multiples, neighbours of multiples, random mantissas over 1900 binades): no mismatch.  Valid
// for normal quotients, which populations always are (0, Inf, NaN behave like the division).
template <int C> __device__ __forceinline__ double divc(double x)
{
    constexpr double r = 1.0 / C;
    const double q0 = __dmul_rn(x, r);
    const double rem = __fma_rn(-(double)C, q0, x);
    return __fma_rn(rem, r, q0);
}

// AB:509-531: the moment transform is written for I-ordered input but receives k-ordered arrays
__device__ __forceinline__ void convert(const double IN[9], double OUT[9])
{
    OUT[0] = IN[0] + IN[1] + IN[2] + IN[3] + IN[4] + IN[5] + IN[6] + IN[7] + IN[8];
    OUT[1] = -IN[1] - IN[2] - IN[3] - IN[4] + 2 * (IN[5] + IN[6] + IN[7] + IN[8]) - 4 * IN[0];
    OUT[2] = (IN[5] + IN[6] + IN[7] + IN[8]) - 2 * (IN[1] + IN[2] + IN[3] + IN[4]) + 4 * IN[0];
    OUT[3] = IN[1] - IN[3] + IN[5] - IN[6] - IN[7] + IN[8];
    OUT[4] = IN[5] - IN[6] - IN[7] + IN[8] - 2 * (IN[1] - IN[3]);
    OUT[5] = IN[2] - IN[4] + IN[5] + IN[6] - IN[7] - IN[8];
    OUT[6] = IN[5] + IN[6] - IN[7] - IN[8] - 2 * (IN[2] - IN[4]);
    OUT[7] = IN[1] - IN[2] + IN[3] - IN[4];
    OUT[8] = IN[5] - IN[6] + IN[7] - IN[8];
}
__device__ __forceinline__ void reconvert(const double IN[9], double OUT[9])
{
    // the reference's expressions with every "/ c" evaluated by divc<c> (same correctly rounded value) and the repeated
    // quotients named once (the compiler would do the same CSE on the division form)
    const double C0 = divc<9>(IN[0]), C7 = IN[7] / 4.0, C8 = IN[8] / 4.0;
    const double a36 = divc<36>(IN[1] + 2 * IN[2]), b36 = divc<36>(IN[2] + 2 * IN[1]);
    const double d34 = divc<6>(IN[3] - IN[4]), d56 = divc<6>(IN[5] - IN[6]);
    const double s35 = divc<6>(IN[3] + IN[5]), d35 = divc<6>(IN[3] - IN[5]);
    const double s46 = divc<12>(IN[4] + IN[6]), d46 = divc<12>(IN[4] - IN[6]);
    OUT[0] = C0 - divc<9>(IN[1] - IN[2]);
    OUT[1] = C0 - a36 + d34 + C7;
    OUT[2] = C0 - a36 + d56 - C7;
    OUT[3] = C0 - a36 - d34 + C7;
    OUT[4] = C0 - a36 - d56 - C7;
    OUT[5] = C0 + b36 + s35 + s46 + C8;
    OUT[6] = C0 + b36 - d35 - d46 - C8;
    OUT[7] = C0 + b36 - s35 - s46 + C8;
    OUT[8] = C0 + b36 + d35 + d46 - C8;
}

// Fobj in the reference's padded coordinates (Xp = X+1 in [1, nx], Yp = Y+1 in [0, ny+1]) as the pure function of the
// wall positions that Initialize_Fobj_for_Vessel_Walls tabulates (AB:275-285).  The extrapolated ghost columns
// Xp = 0, nx+1 are never read by the time step and are not provided.
__host__ __device__ inline double fobj(const double *yr1, const double *yr2, const Geo &g, int Xp, int Yp)
{
    const int X = Xp - 1, Y = Yp - 1;
    return (Y <= g.Y0 ? (yr1[X] - g.c) : (yr2[X] - g.c)) / (Y - g.c);
}

// MRT_Collision of one node held in registers (shared by puls_collide and puls_fused: identical operations)
__device__ __forceinline__ void mrt_post(const Par &mp, const double gin[9], double P, double Ux, double Uy, double post[9])
{
    double geq[9], tmp[9], m[9], dpost[9];
    equilibrium_g(mp.Rho0, P, Ux, Uy, geq);
#pragma unroll
    for (int k = 0; k < 9; ++k) tmp[k] = gin[k] - geq[k];
    convert(tmp, m);
#pragma unroll
    for (int q = 0; q < 9; ++q) m[q] *= mp.S[q];
    reconvert(m, dpost);
#pragma unroll
    for (int k = 0; k < 9; ++k) post[k] = gin[k] - dpost[k];
}

// ---- collide ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) puls_collide(const double *__restrict__ A, double *__restrict__ B,
                                                    const uint8_t *__restrict__ flag, const double *__restrict__ P,
                                                    const double *__restrict__ Ux, const double *__restrict__ Uy, Geo g, Par mp)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.nelem || flag[i] == 0) return;
    double gin[9], post[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) gin[k] = A[k * g.nelem + i];
    mrt_post(mp, gin, P[i], Ux[i], Uy[i], post);
#pragma unroll
    for (int k = 0; k < 9; ++k) B[k * g.nelem + i] = post[k];
}

// ---- fused collide + pull stream + moments for INTERIOR fluid nodes ---------------------------------------
// An interior node is a fluid node off the lattice rim (1 <= X <= nx-2, 1 <= Y <= ny-2) whose eight neighbours are all
// fluid: every population it pulls is a plain post-collision value -- no Bouzidi value, no stale solid-node value, no
// Zou/He column, no flat-index spill.  For those nodes (almost all of them) the reference's
//     collide -> [Bouzidi] -> pull -> [Zou/He] -> moments
// is done here in one pass: a CTA owns NT-2 consecutive rows and marches along x with a 3-column shared-memory ring of
// post-collision populations; column x+1 is collided (and written to the out buffer, which the NEXT iteration's
// collision reads un-streamed -- SURVEY.md B.1), then column x is streamed out of the ring and its P, Ux, Uy are
// written.  The streamed populations of an interior node are NOT written to the in buffer: the reference overwrites
// them with the next post-collision values before anything reads them, except (a) a fresh node being seeded from
// nodes >= 3 links away and (b) a lattice download -- both re-create them on demand (interior_streamed below).
// Everything else (solid nodes, fluid nodes next to a solid node, the rim) goes through puls_stream after the Bouzidi
// kernels, exactly as in the unfused path.  P, Ux, Uy are double buffered: the collision of a halo row / column reads
// the old values while its owner may already have written the new ones.
// Traffic per interior node: 9 + 3 doubles read, 9 + 3 written (192 B) instead of 336 B for collide + stream.
template <int NT>
__global__ void __launch_bounds__(NT) puls_fused(const double *__restrict__ A, double *__restrict__ B, const uint8_t *__restrict__ flag,
                                                 const double *__restrict__ P, const double *__restrict__ Ux,
                                                 const double *__restrict__ Uy, double *__restrict__ Pn, double *__restrict__ Uxn,
                                                 double *__restrict__ Uyn, uint8_t *__restrict__ intr, Geo g, Par mp, int xchunk)
{
    __shared__ double R[3][9][NT];
    __shared__ uint8_t F[3][NT];
    const int tid = threadIdx.x;
    const int Y = (int)blockIdx.x * (NT - 2) - 1 + tid;          // rows tid = 0 and NT-1 are halo rows
    const bool row_ok = Y >= 0 && Y < g.ny;
    const bool own = tid >= 1 && tid < NT - 1 && row_ok;
    const int xa = blockIdx.y * xchunk, xb = min(g.nx, xa + xchunk);
    auto slot_of = [](int X) { return (X + 3) % 3; };
    // the 12 inputs of a column's collision are loaded ONE COLUMN AHEAD into registers, so a thread always has loads in
    // flight while it computes, waits at the barriers and streams (the kernel is bound by memory-level parallelism, not
    // by instruction issue: without the prefetch it moved 3.9 TB/s)
    // The node mask decides whether a node has inputs at all, so it runs TWO columns ahead: with the mask loaded in front of the
    // populations it guards, every column paid two memory latencies in a row (ncu at N = 1024: 23 % of the stall samples on the
    // compare behind the mask load, 19 % on the first use of the prefetched registers).
    double nin[12];
    uint8_t nfl = 0, nnfl = 0;
    auto fetch_flag = [&](int X) {
        nnfl = 0;
        if (row_ok && X >= 0 && X < g.nx) nnfl = flag[Y + (long long)g.ny * X];
    };
    auto fetch_col = [&](int X) {      // the mask of column X is in nnfl (fetch_flag(X) one call earlier)
        nfl = nnfl;
        fetch_flag(X + 1);
        if (!nfl) return;
        const long long i = Y + (long long)g.ny * X;
#pragma unroll
        for (int k = 0; k < 9; ++k) nin[k] = A[k * g.nelem + i];
        nin[9] = P[i]; nin[10] = Ux[i]; nin[11] = Uy[i];
    };
    // collides column X from the prefetched registers, then prefetches column X + 1
    auto collide_col = [&](int X, bool write, bool more) {
        const int s = slot_of(X);
        const uint8_t fl = nfl;
        double gin[9], post[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) gin[k] = nin[k];
        const double p0 = nin[9], u0 = nin[10], v0 = nin[11];
        if (more) fetch_col(X + 1);
        F[s][tid] = fl;
        if (!fl) return;
        mrt_post(mp, gin, p0, u0, v0, post);
#pragma unroll
        for (int k = 0; k < 9; ++k) R[s][k][tid] = post[k];
        if (write && own) {
            const long long i = Y + (long long)g.ny * X;
#pragma unroll
            for (int k = 0; k < 9; ++k) B[k * g.nelem + i] = post[k];
        }
    };
    fetch_flag(xa - 1);
    fetch_col(xa - 1);
    collide_col(xa - 1, false, true);
    collide_col(xa, true, true);
    for (int X = xa; X < xb; ++X) {
        collide_col(X + 1, X + 1 < xb, X + 1 < xb);
        __syncthreads();
        if (own) {
            const int sm = slot_of(X - 1), s0 = slot_of(X), sp = slot_of(X + 1);
            const bool interior = X >= 1 && X <= g.nx - 2 && Y >= 1 && Y <= g.ny - 2 &&
                                  (F[sm][tid - 1] & F[sm][tid] & F[sm][tid + 1] & F[s0][tid - 1] & F[s0][tid] & F[s0][tid + 1] &
                                   F[sp][tid - 1] & F[sp][tid] & F[sp][tid + 1]) != 0;
            intr[Y + (long long)g.ny * X] = interior ? 1 : 0;      // puls_stream skips these nodes (one coalesced byte per node)
            if (interior) {
                double gk[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) gk[k] = R[ckx(k) > 0 ? sm : (ckx(k) < 0 ? sp : s0)][k][tid - cky(k)];   // source (X - cx, Y - cy)
                double pp = 0.0, ux = 0.0, uy = 0.0;
#pragma unroll
                for (int k = 0; k < 9; ++k) pp += gk[k];
#pragma unroll
                for (int k = 1; k < 9; ++k) { ux += gk[k] * ckx(k); uy += gk[k] * cky(k); }
                const long long i = Y + (long long)g.ny * X;
                Pn[i] = pp;
                Uxn[i] = (mp.Rho0 == 1.0) ? 3.0 * ux : 3.0 * ux / mp.Rho0;
                Uyn[i] = (mp.Rho0 == 1.0) ? 3.0 * uy : 3.0 * uy / mp.Rho0;
            }
        }
        __syncthreads();
    }
}

// ---- puls_fused with its inputs staged by TMA (opt-in, CLBM_PULS_TMA) ------------------------------------------------------
// An experiment with a negative result, kept because it is bit-exact and documents where the kernel stands.  ncu at N = 1024 shows
// 44 % of puls_fused's stall samples as long-scoreboard waits, which is what TMA staging cured in the D2Q9 Shan-Chen kernel.  Not
// here: measured 361 us (2 stages) / 351 us (3 stages) against 351 us -- puls_fused already moves 1.05 GB + 0.96 GB per launch,
// 5.75 TB/s, 0.88 of the measured HBM peak; the waits are the DRAM queue, not exposed latency.  What is left of the iteration is the
// 100 us of one-thread-per-column wall kernels behind it.
// The 12 input columns of a tile -- 9 populations, P, Ux, Uy -- arrive
// as four cp.async.bulk.tensor boxes per column in an NS-stage shared-memory pipeline with mbarrier completion: no load instruction,
// no address arithmetic and no registers are spent on data in flight, and NS - 1 whole columns are always on their way.  The box starts
// one row in front of the tile's first halo row (an even row: the innermost TMA coordinate has to be 16-byte aligned) and is NT + 2
// rows tall; rows and columns outside the lattice are zero-filled by the TMA unit and never used (their mask is 0).  Solid nodes are
// fetched too (a box cannot skip them).  Arithmetic, ring and stores are those of puls_fused: the results are bit-identical.
template <int NT, int NS>
struct PulsTmaCfg {
    static constexpr int BY = NT + 2;
    static constexpr int A_BYTES = 9 * BY * 8, F_BYTES = BY * 8;
    static constexpr int A_PITCH = ((A_BYTES + 127) / 128) * 128, F_PITCH = ((F_BYTES + 127) / 128) * 128;
    static constexpr int STAGE_BYTES = A_PITCH + 3 * F_PITCH;
    static constexpr int TX = A_BYTES + 3 * F_BYTES;
    static constexpr int OFF_R = NS * STAGE_BYTES;            // R[3][9][NT]
    static constexpr int OFF_F = OFF_R + 3 * 9 * NT * 8;      // F[3][NT]
    static constexpr int OFF_BAR = ((OFF_F + 3 * NT + 15) / 16) * 16;
    static constexpr int SMEM = OFF_BAR + NS * 8;
    static_assert(NT % 2 == 0 && BY <= 256, "even tile height (16-byte rows), one TMA box dimension");
};

template <int NT, int NS>
__global__ void __launch_bounds__(NT) puls_fused_tma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmP,
                                                     const __grid_constant__ CUtensorMap tmUx, const __grid_constant__ CUtensorMap tmUy,
                                                     double *__restrict__ B, const uint8_t *__restrict__ flag, double *__restrict__ Pn,
                                                     double *__restrict__ Uxn, double *__restrict__ Uyn, uint8_t *__restrict__ intr, Geo g, Par mp,
                                                     int xchunk)
{
    using C = PulsTmaCfg<NT, NS>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t stage_a = smem_u32(smem_raw);
    double (*R)[9][NT] = reinterpret_cast<double (*)[9][NT]>(smem_raw + C::OFF_R);
    uint8_t (*F)[NT] = reinterpret_cast<uint8_t (*)[NT]>(smem_raw + C::OFF_F);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + C::OFF_BAR);
    const int tid = threadIdx.x;
    const int Y = (int)blockIdx.x * (NT - 2) - 1 + tid;          // rows tid = 0 and NT-1 are halo rows
    const int Ys = (int)blockIdx.x * (NT - 2) - 2;               // first row of the boxes: thread tid reads box row tid + 1
    const bool row_ok = Y >= 0 && Y < g.ny;
    const bool own = tid >= 1 && tid < NT - 1 && row_ok;
    const int xa = blockIdx.y * xchunk, xb = min(g.nx, xa + xchunk);
    const int ncols = xb - xa + 2;                               // r = 0 .. ncols-1 enumerates the columns xa-1 .. xb
    auto slot_of = [](int X) { return (X + 3) % 3; };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(&mbar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int r) {                                    // thread 0: the four boxes of column r into stage r % NS
        const int X = xa - 1 + r;
        uint64_t *bar = &mbar[r % NS];
        const uint32_t st = stage_a + (r % NS) * C::STAGE_BYTES;
        mbar_expect_tx(bar, (uint32_t)C::TX);
        tma_load_3d(st, &tmA, bar, Ys, X, 0);
        tma_load_2d(st + C::A_PITCH, &tmP, bar, Ys, X);
        tma_load_2d(st + C::A_PITCH + C::F_PITCH, &tmUx, bar, Ys, X);
        tma_load_2d(st + C::A_PITCH + 2 * C::F_PITCH, &tmUy, bar, Ys, X);
    };
    // stage r % NS is free once every thread has its column-r inputs in registers: the caller has a barrier behind collide_col
    auto refill = [&](int r) {
        if (tid == 0 && r + NS < ncols) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(r + NS);
        }
    };
    auto flag_of = [&](int X) -> uint8_t { return (row_ok && X >= 0 && X < g.nx) ? flag[Y + (long long)g.ny * X] : (uint8_t)0; };
    uint8_t f0, f1;          // node masks of the next two columns to be collided: plain loads, two columns ahead of their use
    auto collide_col = [&](int r, bool write) {
        const int X = xa - 1 + r;
        const int s = slot_of(X);
        const uint8_t fl = f0;
        f0 = f1;
        f1 = (r + 2 < ncols) ? flag_of(X + 2) : (uint8_t)0;
        F[s][tid] = fl;
        mbar_wait(&mbar[r % NS], (uint32_t)((r / NS) & 1));
        if (!fl) return;
        const uint32_t st = stage_a + (r % NS) * C::STAGE_BYTES + (tid + 1) * 8;
        double gin[9], post[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) gin[k] = lds_f64(st + k * (C::BY * 8));
        const double p0 = lds_f64(st + C::A_PITCH), u0 = lds_f64(st + C::A_PITCH + C::F_PITCH), v0 = lds_f64(st + C::A_PITCH + 2 * C::F_PITCH);
        mrt_post(mp, gin, p0, u0, v0, post);
#pragma unroll
        for (int k = 0; k < 9; ++k) R[s][k][tid] = post[k];
        if (write && own) {
            const long long i = Y + (long long)g.ny * X;
#pragma unroll
            for (int k = 0; k < 9; ++k) B[k * g.nelem + i] = post[k];
        }
    };
    if (tid == 0) {
#pragma unroll
        for (int r = 0; r < NS; ++r)
            if (r < ncols) issue(r);
    }
    f0 = flag_of(xa - 1);
    f1 = flag_of(xa);
    collide_col(0, false);
    __syncthreads();
    refill(0);
    collide_col(1, true);
    __syncthreads();
    refill(1);
    for (int X = xa; X < xb; ++X) {
        const int r = X - xa + 2;                                // column X + 1
        collide_col(r, X + 1 < xb);
        __syncthreads();
        refill(r);
        if (own) {
            const int sm = slot_of(X - 1), s0 = slot_of(X), sp = slot_of(X + 1);
            const bool interior = X >= 1 && X <= g.nx - 2 && Y >= 1 && Y <= g.ny - 2 &&
                                  (F[sm][tid - 1] & F[sm][tid] & F[sm][tid + 1] & F[s0][tid - 1] & F[s0][tid] & F[s0][tid + 1] &
                                   F[sp][tid - 1] & F[sp][tid] & F[sp][tid + 1]) != 0;
            intr[Y + (long long)g.ny * X] = interior ? 1 : 0;
            if (interior) {
                double gk[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) gk[k] = R[ckx(k) > 0 ? sm : (ckx(k) < 0 ? sp : s0)][k][tid - cky(k)];   // source (X - cx, Y - cy)
                double pp = 0.0, ux = 0.0, uy = 0.0;
#pragma unroll
                for (int k = 0; k < 9; ++k) pp += gk[k];
#pragma unroll
                for (int k = 1; k < 9; ++k) { ux += gk[k] * ckx(k); uy += gk[k] * cky(k); }
                const long long i = Y + (long long)g.ny * X;
                Pn[i] = pp;
                Uxn[i] = (mp.Rho0 == 1.0) ? 3.0 * ux : 3.0 * ux / mp.Rho0;
                Uyn[i] = (mp.Rho0 == 1.0) ? 3.0 * uy : 3.0 * uy / mp.Rho0;
            }
        }
        __syncthreads();
    }
}

// streamed ("gin" after Streaming) population k of a node that was interior when the last step ran; B = that step's out buffer
__device__ __forceinline__ double interior_streamed(const double *__restrict__ B, const Geo &g, int X, int Y, int k)
{
    return B[k * g.nelem + (Y - cky(k)) + (long long)g.ny * (X - ckx(k))];
}
// the same predicate from the wall positions the step ran with (the mask array has moved on by the time it is asked)
__device__ inline bool interior_by_walls(const double *yr1o, const double *yr2o, const Geo &g, int X, int Y)
{
    if (X < 1 || X > g.nx - 2 || Y < 1 || Y > g.ny - 2) return false;
    for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy)
            if (fobj(yr1o, yr2o, g, X + dx + 1, Y + dy + 1) < 1.0) return false;
    return true;
}
// writes the skipped streamed populations of the last step's interior nodes into its in buffer (lattice download)
__global__ void __launch_bounds__(256) puls_materialise(double *__restrict__ A, const double *__restrict__ B, const double *__restrict__ yr1o,
                                                        const double *__restrict__ yr2o, Geo g)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.nelem) return;
    const int X = (int)(i / g.ny), Y = (int)(i % g.ny);
    if (!interior_by_walls(yr1o, yr2o, g, X, Y)) return;
#pragma unroll
    for (int k = 0; k < 9; ++k) A[k * g.nelem + i] = interior_streamed(B, g, X, Y, k);
}

// ---- Bouzidi ---------------------------------------------------------------------------------------
// AB:288-290
__device__ __forceinline__ double find_delta(int mA, double mB, double Y1)
{
    double D = 1.0 - fabs(Y1 / (mA - mB));
    if (D < 0) D = 0;
    return D;
}

// AB:553-601 for one border node
__device__ void bouzidi_node(double *__restrict__ B, const double *yr1, const double *yr2, const Geo &g, int X, int Y, const double Dl[8])
{
    const int nx = g.nx, ny = g.ny;
    auto indom = [&](int Xp, int Yp) { return Xp >= 0 && Xp < nx && Yp >= 0 && Yp < ny; };
    if (!indom(X, Y)) return;
#pragma unroll
    for (int I = 1; I <= 8; ++I) {
        const double D = Dl[I - 1];
        if (D >= 1.0) continue;
        const int kI = kfromI(I), kJ = kfromI(jbI(I));
        const int X1 = X + exI(I), Y1 = Y + eyI(I);
        int X2 = X1 + exI(I), Y2 = Y1 + eyI(I);
        int X3 = X2 + exI(I), Y3 = Y2 + eyI(I);
        if (!indom(X1, Y1)) continue;
        if (!indom(X2, Y2)) { X2 = X1; Y2 = Y1; }
        if (!indom(X3, Y3)) { X3 = X1; Y3 = Y1; }
        if (!indom(X3, Y3)) { X3 = X2; Y3 = Y2; }
        if (fobj(yr1, yr2, g, X2 + 1, Y2 + 1) < 1) { X2 = X1; Y2 = Y1; }
        if (fobj(yr1, yr2, g, X3 + 1, Y3 + 1) < 1) { X3 = X2; Y3 = Y2; }
        const long long b = Y + (long long)ny * X, n1 = Y1 + (long long)ny * X1, n2 = Y2 + (long long)ny * X2, n3 = Y3 + (long long)ny * X3;
        if (D < 0.5) {
            B[kI * g.nelem + b] = B[kJ * g.nelem + n1] * (1 + 2 * D) * D + B[kJ * g.nelem + n2] * (1 - 2 * D) * (1 + 2 * D) -
                                  B[kJ * g.nelem + n3] * (1 - 2 * D) * D;
        } else {
            B[kI * g.nelem + b] = (B[kJ * g.nelem + n1] - B[kI * g.nelem + n1] * (1 - 2 * D) * (1 + 2 * D) + B[kI * g.nelem + n2] * (1 - 2 * D) * D) /
                                  (D * (1 + 2 * D));
        }
    }
}

__device__ __forceinline__ void d_reset(double D[8])
{
#pragma unroll
    for (int i = 0; i < 8; ++i) D[i] = 2;
}

// row of the bottom / top border node of column X (AB:299-300, :344-345)
__device__ __forceinline__ int border_row(const double *yr1, const double *yr2, const Geo &g, int X, int top)
{
    int Y = top ? (int)ceil(yr2[X]) : (int)floor(yr1[X]);
    if (fobj(yr1, yr2, g, X + 1, Y + 1) >= 1) Y += top ? 1 : -1;
    return Y;
}

// thread X of wall `top`: the border entries the reference's sweep appends while visiting column X (a transitional
// entry when the border row changes between X-1 and X, then the column's own entry), applied immediately.
// Within one wall's list no two entries share a (node, link) and no entry reads a slot another entry of the same
// list writes, so the order inside a list does not matter.  Across the lists it does where the vessel is pinched to
// less than one fluid row (the bottom entry's first link node is then the top border node and vice versa): the
// reference applies Borders1 before Borders2 (AB:543-551), hence one launch per wall, bottom first.
__global__ void __launch_bounds__(128) puls_bouzidi(double *__restrict__ B, const double *__restrict__ yr1, const double *__restrict__ yr2, Geo g, int top)
{
    const int X = blockIdx.x * blockDim.x + threadIdx.x, nx = g.nx;
    if (X >= nx) return;
    auto FO = [&](int Xp, int Yp) { return fobj(yr1, yr2, g, Xp, Yp); };
    double D[8];
    const int Yx = border_row(yr1, yr2, g, X, top);
    if (!top) {
        if (X >= 1) {
            const int Y = border_row(yr1, yr2, g, X - 1, 0);
            if (Yx != Y) {
                d_reset(D);
                if (Yx > Y) { D[5] = find_delta(-1, yr1[X] - yr1[X - 1], yr1[X] - Y); bouzidi_node(B, yr1, yr2, g, X, Y, D); }
                else { D[4] = find_delta(1, yr1[X] - yr1[X - 1], yr1[X - 1] - Yx); bouzidi_node(B, yr1, yr2, g, X - 1, Yx, D); }
            }
        }
        d_reset(D);
        if (X < nx - 1 && FO(X + 2, Yx + 1) >= 1) D[0] = find_delta(0, yr1[X + 1] - yr1[X], yr1[X] - Yx);
        D[1] = 1 - (yr1[X] - Yx);
        if (X >= 1 && FO(X, Yx + 1) >= 1) D[2] = find_delta(0, yr1[X] - yr1[X - 1], yr1[X] - Yx);
        if (X < nx - 1 && FO(X + 2, Yx + 2) >= 1) D[4] = find_delta(1, yr1[X + 1] - yr1[X], yr1[X] - Yx);
        if (X >= 1 && FO(X, Yx + 2) >= 1) D[5] = find_delta(-1, yr1[X] - yr1[X - 1], yr1[X] - Yx);
        bouzidi_node(B, yr1, yr2, g, X, Yx, D);
    } else {
        if (X >= 1) {
            const int Yp = border_row(yr1, yr2, g, X - 1, 1);
            if (Yx != Yp) {
                d_reset(D);
                if (Yx > Yp) { D[7] = find_delta(-1, yr2[X] - yr2[X - 1], yr2[X - 1] - Yx); bouzidi_node(B, yr1, yr2, g, X - 1, Yx, D); }
                else { D[6] = find_delta(1, yr2[X] - yr2[X - 1], yr2[X] - Yp); bouzidi_node(B, yr1, yr2, g, X, Yp, D); }
            }
        }
        d_reset(D);
        if (X < nx - 1 && FO(X + 2, Yx + 1) >= 1) D[0] = find_delta(0, yr2[X + 1] - yr2[X], yr2[X] - Yx);
        if (X >= 1 && FO(X, Yx + 1) >= 1) D[2] = find_delta(0, yr2[X] - yr2[X - 1], yr2[X] - Yx);
        D[3] = 1 - (Yx - yr2[X]);
        if (X >= 1 && FO(X, Yx) >= 1) D[6] = find_delta(1, yr2[X] - yr2[X - 1], yr2[X] - Yx);
        if (X < nx - 1 && FO(X + 2, Yx) >= 1) D[7] = find_delta(-1, yr2[X + 1] - yr2[X], yr2[X] - Yx);
        bouzidi_node(B, yr1, yr2, g, X, Yx, D);
    }
}

// ---- pull streaming + Zou/He + macroscopic ----------------------------------------------------------
// `lat` is the whole allocation; bin/bout are the offsets of the buffers the reference calls gin / gout.
__device__ __forceinline__ void puls_stream_node(long long i, double *__restrict__ lat, long long bin, long long bout, const uint8_t *__restrict__ flag,
                                                 double *__restrict__ P, double *__restrict__ Ux, double *__restrict__ Uy,
                                                 const double *__restrict__ yr1, const double *__restrict__ yr2, const Geo &g, const Par &mp,
                                                 double Pin, double Pout)
{
    const int X = (int)(i / g.ny), Y = (int)(i % g.ny);
    const double *B = lat + bout;
    double gk[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int XX = (X - ckx(k) + g.nx) % g.nx;
        const int YY = Y - cky(k);                      // wrapY is the identity: flat-index spill (B.3)
        const long long src = (long long)YY + (long long)g.ny * XX;
        if (k == 8 && src == g.nelem && bout + 9 * g.nelem == bin) {
            // the spill of (0, ny-1, k=8) lands on gin(node 0, k=0), which the reference's X-major sweep has
            // already overwritten with its streamed value gout((1,0), k=0)
            gk[k] = B[0 * g.nelem + (long long)g.ny * (1 % g.nx)];
        } else {
            gk[k] = B[k * g.nelem + src];
        }
    }
    // Zou/He pressure inlet (AB:618-642) / outlet (AB:644-669), I-ordered names
    const double Rho0 = mp.Rho0;
    if (X == 0) {
        int ylo = (int)ceil(yr1[0] - 0.01), yhi = (int)floor(yr2[0] + 0.01);
        if (ylo < 0) ylo = 0;
        if (yhi > g.ny - 1) yhi = g.ny - 1;
        if (Y >= ylo && Y <= yhi) {
            const double g0 = gk[kfromI(0)], g2 = gk[kfromI(2)], g3 = gk[kfromI(3)], g4 = gk[kfromI(4)], g6 = gk[kfromI(6)], g7 = gk[kfromI(7)];
            double Uin = Pin - g0 - g2 - 2 * g3 - g4 - 2 * g6 - 2 * g7;
            Uin = Uin * 3.0 / Rho0;
            gk[kfromI(1)] = g3 + 2.0 * Rho0 / 9.0 * Uin;
            gk[kfromI(5)] = Rho0 / 18.0 * Uin - 0.5 * (g2 - g4) + g7;
            gk[kfromI(8)] = Rho0 / 18.0 * Uin + 0.5 * (g2 - g4) + g6;
        }
    }
    if (X == g.nx - 1) {       // after the inlet, as in the reference (matters only for nx == 1)
        int ylo = (int)ceil(yr1[g.nx - 1] - 0.01), yhi = (int)floor(yr2[g.nx - 1] + 0.01);
        if (ylo < 0) ylo = 0;
        if (yhi > g.ny - 1) yhi = g.ny - 1;
        if (Y >= ylo && Y <= yhi) {
            const double g0 = gk[kfromI(0)], g1 = gk[kfromI(1)], g2 = gk[kfromI(2)], g4 = gk[kfromI(4)], g5 = gk[kfromI(5)], g8 = gk[kfromI(8)];
            double Uout = g0 + 2 * g1 + g2 + g4 + 2 * g5 + 2 * g8 - Pout;
            Uout = Uout * 3.0 / Rho0;
            gk[kfromI(3)] = g1 - 2.0 * Rho0 / 9.0 * Uout;
            gk[kfromI(6)] = -Rho0 / 18.0 * Uout - 0.5 * (g2 - g4) + g8;
            gk[kfromI(7)] = -Rho0 / 18.0 * Uout + 0.5 * (g2 - g4) + g5;
        }
    }
    double *A = lat + bin;
#pragma unroll
    for (int k = 0; k < 9; ++k) A[k * g.nelem + i] = gk[k];
    // Macroscopic_Properties_g (AB:216-230): the velocity sums start at k = 1 (as written)
    if (flag[i] == 0) {
        P[i] = 0.0; Ux[i] = 0.0; Uy[i] = 0.0;
    } else {
        double pp = 0.0, ux = 0.0, uy = 0.0;
#pragma unroll
        for (int k = 0; k < 9; ++k) pp += gk[k];
#pragma unroll
        for (int k = 1; k < 9; ++k) { ux += gk[k] * ckx(k); uy += gk[k] * cky(k); }
        P[i] = pp;
        Ux[i] = (Rho0 == 1.0) ? 3.0 * ux : 3.0 * ux / Rho0;
        Uy[i] = (Rho0 == 1.0) ? 3.0 * uy : 3.0 * uy / Rho0;
    }
}

// One thread looks at PULS_SCAN consecutive nodes.  With the fused interior pass almost every node is marked `intr` (done by
// puls_fused) and a thread has nothing to do: one 16-byte load of the marks instead of sixteen threads that each load a byte
// and leave (57 us for the 10.5 M nodes of N = 1024, of which 15 000 are not interior).
template <int PULS_SCAN>
__global__ void __launch_bounds__(256) puls_stream(double *__restrict__ lat, long long bin, long long bout, const uint8_t *__restrict__ flag,
                                                   double *__restrict__ P, double *__restrict__ Ux, double *__restrict__ Uy,
                                                   const double *__restrict__ yr1, const double *__restrict__ yr2, Geo g, Par mp,
                                                   double Pin, double Pout, const uint8_t *__restrict__ intr)
{
    const long long block_base = (long long)blockIdx.x * blockDim.x * PULS_SCAN;
    const long long base = block_base + (long long)threadIdx.x * PULS_SCAN;
    if (PULS_SCAN == 1) {   // the two-pass form (no interior marks): one node per thread, coalesced
        if (base >= g.nelem) return;
        if (intr && intr[base]) return;
        puls_stream_node(base, lat, bin, bout, flag, P, Ux, Uy, yr1, yr2, g, mp, Pin, Pout);
        return;
    }
    unsigned done = 0;      // bit j: node base + j was handled by puls_fused
    if (intr) {
        if (base + PULS_SCAN <= g.nelem) {
            const uint4 m = *reinterpret_cast<const uint4 *>(intr + base);   // cudaMalloc alignment, base a multiple of 16
            const unsigned w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if ((w[q] >> (8 * b)) & 0xffu) done |= 1u << (4 * q + b);
        } else {
            for (int j = 0; j < PULS_SCAN && base + j < g.nelem; ++j)
                if (intr[base + j]) done |= 1u << j;
        }
    }
    // the nodes that are left (the two rim columns, the bands along the walls) are queued in shared memory and shared out over
    // the block: a thread that walked its own 16 nodes one after the other made the launch slower than the byte-per-thread form
    __shared__ int todo[256 * PULS_SCAN];
    __shared__ int ntodo;
    if (threadIdx.x == 0) ntodo = 0;
    __syncthreads();
    if (base < g.nelem && done != 0xffffu) {
        int n = 0;
        for (int j = 0; j < PULS_SCAN; ++j)
            if (!((done >> j) & 1u) && base + j < g.nelem) ++n;
        int at = atomicAdd(&ntodo, n);
        for (int j = 0; j < PULS_SCAN; ++j)
            if (!((done >> j) & 1u) && base + j < g.nelem) todo[at++] = (int)(base + j - block_base);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < ntodo; q += blockDim.x)
        puls_stream_node(block_base + todo[q], lat, bin, bout, flag, P, Ux, Uy, yr1, yr2, g, mp, Pin, Pout);
}

// ---- wall motion (AB:243-272) ---------------------------------------------------------------------
__global__ void __launch_bounds__(128) puls_walls(const double *__restrict__ P, double *__restrict__ yr1, double *__restrict__ yr2,
                                                  double *__restrict__ yr1o, double *__restrict__ yr2o, Geo g, Par mp)
{
    const int X = blockIdx.x * blockDim.x + threadIdx.x;
    if (X >= g.nx) return;
    const double Yw1 = 0.0, Yw2 = (double)(g.ny - 1), cap = 0.25;
    {
        const double Ps = P[g.Y0 + (long long)g.ny * X] - mp.p_tissue;
        const double target = (Yw1 + 0.5) - Ps / mp.alpha;
        double d = target - yr1[X];
        if (d > cap) d = cap;
        if (d < -cap) d = -cap;
        yr1o[X] = yr1[X];
        yr1[X] = yr1[X] + d;
    }
    {
        const double Ps = P[g.Y0 + 1 + (long long)g.ny * X] - mp.p_tissue;
        const double target = (Yw2 - 0.5) + Ps / mp.alpha;
        double d = target - yr2[X];
        if (d > cap) d = cap;
        if (d < -cap) d = -cap;
        yr2o[X] = yr2[X];
        yr2[X] = yr2[X] + d;
    }
}

// ---- geometry update + fresh nodes (AB:384-498) ------------------------------------------------------
// one thread per column X: re-evaluates the mask where the wall moved and fills the nodes that turned fluid.
// A fill reads populations of nodes that were fluid before the move (weight (int)Fold != 0) or, in the end
// columns, the node next to it in the same column -- never a node another thread fills -- so columns are
// independent.  The exception is a fresh node with no old-fluid node around it (a closed stretch of the vessel
// opening up): the reference seeds it from the nearest nodes that are fluid NOW (Seed_From_Nearest_Fluid,
// AB:418-458), which may be other fresh nodes, filled or not yet filled depending on the X-major/Y-minor sweep
// order.  Those nodes are only queued here (with the pre-fill populations of every fresh node saved in `pre`) and
// handled in sweep order by puls_seed.
constexpr int PRE_ROWS = 8;     // a wall moves <= 0.25 per step: the re-evaluated window of a half column is <= 7 rows
__device__ __forceinline__ long long pre_slot(int X, int Y, const Geo &g) { return (((long long)X * 2 + (Y > g.Y0)) * PRE_ROWS + (Y & (PRE_ROWS - 1))) * 9; }

__global__ void __launch_bounds__(128) puls_fobj(double *__restrict__ A, uint8_t *__restrict__ flag, double *__restrict__ P,
                                                 double *__restrict__ Ux, double *__restrict__ Uy, const double *__restrict__ yr1,
                                                 const double *__restrict__ yr2, const double *__restrict__ yr1o,
                                                 const double *__restrict__ yr2o, Geo g, Par mp, double *__restrict__ pre,
                                                 int *__restrict__ seed_count, int *__restrict__ seed_list, int seed_cap,
                                                 int *__restrict__ err)
{
    const int X = blockIdx.x * blockDim.x + threadIdx.x;
    if (X >= g.nx) return;
    const int nx = g.nx, ny = g.ny;
    auto clampi = [](double v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : (int)v); };
    for (int half = 0; half < 2; ++half) {
        const double a = half ? yr2o[X] : yr1o[X], b = half ? yr2[X] : yr1[X];
        int ylo = clampi(floor(fmin(a, b)) - 2.0, 0, ny - 1), yhi = clampi(ceil(fmax(a, b)) + 2.0, 0, ny - 1);
        if (!(a == a) || !(b == b)) { ylo = 0; yhi = ny - 1; }
        if (half == 0) { if (yhi > g.Y0) yhi = g.Y0; }
        else { if (ylo < g.Y0 + 1) ylo = g.Y0 + 1; }
        for (int Y = ylo; Y <= yhi; ++Y) {
            const double Fo = fobj(yr1o, yr2o, g, X + 1, Y + 1), Fn = fobj(yr1, yr2, g, X + 1, Y + 1);
            const long long id = Y + (long long)ny * X;
            flag[id] = Fn < 1.0 ? 0 : 1;
            if (!(Fo < 1 && Fn >= 1)) continue;
            if (yhi - ylo >= PRE_ROWS) { atomicExch(err, 2); continue; }
            {
                const long long ps = pre_slot(X, Y, g);
#pragma unroll
                for (int k = 0; k < 9; ++k) pre[ps + k] = A[k * g.nelem + id];
            }
            // Fill_Fluid_Node (AB:460-487)
            if (X == 0 || X == nx - 1) {
                const int Ys = (Y < g.Y0) ? Y + 1 : Y - 1;
                const long long is = Ys + (long long)ny * X;
#pragma unroll
                for (int I = 0; I < 9; ++I) A[kfromI(I) * g.nelem + id] = A[kfromI(I) * g.nelem + is];
            } else {
                int Ff[3][3], Sum = 0;
#pragma unroll
                for (int i = -1; i <= 1; ++i)
#pragma unroll
                    for (int j = -1; j <= 1; ++j) {
                        // Fold rows Yp = Y+1+j in [0, ny+1] exist in the reference's padded table
                        Ff[i + 1][j + 1] = (int)fobj(yr1o, yr2o, g, X + 1 + i, Y + 1 + j);
                        Sum += Ff[i + 1][j + 1];
                    }
                if (Sum == 0) {      // Seed_From_Nearest_Fluid: order dependent, queued for puls_seed
                    const int slot = atomicAdd(seed_count, 1);
                    if (slot < seed_cap) seed_list[slot] = (int)id;
                    else atomicExch(err, 1);
                    continue;
                }
                auto W = [&](int dx, int dy, int k) {   // population of neighbour (X+dx, Y+dy), skipped when its weight is 0
                    const int w = Ff[dx + 1][dy + 1];
                    return w == 0 ? 0.0 : A[k * g.nelem + (Y + dy) + (long long)ny * (X + dx)] * w;
                };
#pragma unroll
                for (int I = 0; I < 9; ++I) {
                    if (Ff[1 - exI(I)][1 - eyI(I)] == 1) continue;
                    const int k = kfromI(I);
                    double acc = 0.0;
                    acc += W(-1, -1, k);
                    acc += W(0, -1, k);
                    acc += W(1, -1, k);
                    acc += W(-1, 0, k);
                    acc += W(1, 0, k);
                    acc += W(-1, 1, k);
                    acc += W(0, 1, k);
                    acc += W(1, 1, k);
                    A[k * g.nelem + id] = acc / (double)Sum;
                }
            }
            // Fresh_Macroscopic_Values (AB:489-498)
            double pp = 0, ux = 0, uy = 0;
#pragma unroll
            for (int I = 0; I < 9; ++I) pp += A[kfromI(I) * g.nelem + id];
#pragma unroll
            for (int I = 1; I < 9; ++I) { ux += A[kfromI(I) * g.nelem + id] * exI(I); uy += A[kfromI(I) * g.nelem + id] * eyI(I); }
            P[id] = pp;
            Ux[id] = 3 * ux / mp.Rho0;
            Uy[id] = 3 * uy / mp.Rho0;
        }
    }
}

// Seed_From_Nearest_Fluid (AB:418-458) + Fresh_Macroscopic_Values for the queued nodes, in the reference's sweep
// order (ascending node index).  One warp: lane k < 9 accumulates population k, the sources are visited in the
// reference's order.  A source that is itself a fresh node LATER in the sweep is read with its pre-fill populations.
__global__ void __launch_bounds__(32) puls_seed(double *__restrict__ A, const uint8_t *__restrict__ flag, double *__restrict__ P,
                                                double *__restrict__ Ux, double *__restrict__ Uy, const double *__restrict__ yr1,
                                                const double *__restrict__ yr2, const double *__restrict__ yr1o,
                                                const double *__restrict__ yr2o, Geo g, Par mp, const double *__restrict__ pre,
                                                int *__restrict__ seed_count, int *__restrict__ seed_list, int seed_cap,
                                                const double *__restrict__ B, int fused)
{
    int n = *seed_count;
    if (n == 0) return;
    if (n > seed_cap) n = seed_cap;
    const int lane = threadIdx.x, nx = g.nx, ny = g.ny;
    if (lane == 0) {
        for (int a = 1; a < n; ++a) {     // insertion sort: the queue is short
            const int v = seed_list[a];
            int b = a - 1;
            while (b >= 0 && seed_list[b] > v) { seed_list[b + 1] = seed_list[b]; --b; }
            seed_list[b + 1] = v;
        }
    }
    __syncwarp();
    for (int s = 0; s < n; ++s) {
        const int id = seed_list[s], X = id / ny, Y = id % ny;
        double acc = 0.0;
        int cnt = 0;
        auto visit = [&](int Xn, int Yn) {
            if (Xn < 0 || Xn >= nx || Yn < 0 || Yn >= ny) return;
            const int idn = Yn + ny * Xn;
            if (flag[idn] == 0) return;
            const bool later_fresh = idn > id && fobj(yr1o, yr2o, g, Xn + 1, Yn + 1) < 1 && fobj(yr1, yr2, g, Xn + 1, Yn + 1) >= 1;
            // fused step: an interior node's streamed populations were never stored -- pull them from the out buffer
            const bool skipped = fused && interior_by_walls(yr1o, yr2o, g, Xn, Yn);
            if (lane < 9)
                acc += later_fresh ? pre[pre_slot(Xn, Yn, g) + lane]
                                   : (skipped ? interior_streamed(B, g, Xn, Yn, lane) : A[lane * g.nelem + idn]);
            ++cnt;
        };
        constexpr int dx[8] = {1, -1, 0, 0, 1, 1, -1, -1}, dy[8] = {0, 0, 1, -1, 1, -1, 1, -1};
#pragma unroll
        for (int q = 0; q < 8; ++q) visit(X + dx[q], Y + dy[q]);
        for (int R = 2; cnt == 0 && R <= 4; ++R)
            for (int sx = -R; sx <= R; ++sx) {
                const int sy_top = R - abs(sx);
                visit(X + sx, Y + sy_top);
                visit(X + sx, Y - sy_top);
            }
        if (lane < 9) {
            if (cnt > 0) {
                A[lane * g.nelem + id] = acc / (double)cnt;
            } else {
                double geq[9];
                equilibrium_g(mp.Rho0, P[id], 0.0, 0.0, geq);
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < 9; ++k) if (k == lane) v = geq[k];
                A[lane * g.nelem + id] = v;
            }
        }
        __syncwarp();
        if (lane == 0) {
            double pp = 0, ux = 0, uy = 0;
#pragma unroll
            for (int I = 0; I < 9; ++I) pp += A[kfromI(I) * g.nelem + id];
#pragma unroll
            for (int I = 1; I < 9; ++I) { ux += A[kfromI(I) * g.nelem + id] * exI(I); uy += A[kfromI(I) * g.nelem + id] * eyI(I); }
            P[id] = pp;
            Ux[id] = 3 * ux / mp.Rho0;
            Uy[id] = 3 * uy / mp.Rho0;
        }
        __syncwarp();
    }
    if (lane == 0) *seed_count = 0;
}

}  // namespace puls
}  // namespace clbm

using namespace clbm;
using namespace clbm::puls;

struct clbm_pulsatile {
    clbm_pulsatile_params prm;
    Geo g;
    Par mp;
    int device, parity, t_iter;
    // derived (AB:147-168)
    double p0_in, p0_out, p_tissue, p_osc, omega;
    int t_beat, t_prop, t_start, t_sever;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    int64_t launches;
    double *lat;          // 2*npop + 1 doubles
    uint8_t *flag;
    double *P, *Ux, *Uy, *yr1, *yr2, *yr1o, *yr2o;
    uint8_t *intr;            // fused step: 1 where puls_fused streamed the node (interior), written every step
    double *P2, *Ux2, *Uy2;   // fused step: the set the running step writes (swapped with P, Ux, Uy afterwards)
    int fused;                // 1: puls_fused + puls_stream on the remaining nodes (default), 0: puls_collide + puls_stream
    int tma;                  // fused step through puls_fused_tma: 0 off (CLBM_PULS_TMA=0, odd N, no driver entry point), else the stage count
    int xchunk_env, nt_env;   // CLBM_PULS_XCHUNK / CLBM_PULS_NT, read once in clbm_pulsatile_create
    int tma_xchunk;           // chunk length of the TMA kernel, chosen at its first launch (0 = not yet)
    struct TmapSlot { const void *base; int rank, rows; CUtensorMap map; };
    std::vector<TmapSlot> tmaps;   // tensor maps of the two lattice buffers and the six field arrays, encoded on first use
    bool in_complete;         // the last step's in buffer holds the streamed populations of ALL nodes (see puls_materialise)
    int *err_dev, *err_host;
    double *pre;          // pre-fill populations of this step's fresh nodes
    int *seed_count, *seed_list, seed_cap;
    bool ktiming;
    int ktiming_cap;
    std::vector<cudaEvent_t> kev;
};

namespace {

void derive_parameters(clbm_pulsatile *c)
{
    // Setup_Simulation_Parameters (AB:147-168)
    const clbm_pulsatile_params &p = c->prm;
    c->t_beat = p.t_beat > 0 ? p.t_beat : (c->g.nx > 1 ? c->g.nx : 1);
    c->omega = 2.0 * 3.141592653589793 / (double)c->t_beat;
    c->p0_in = p.p0_in;
    c->p0_out = p.p0_out;
    if (c->p0_in == 0.0 && c->p0_out == 0.0) { c->p0_in = 0.20; c->p0_out = 0.19; }
    if (p.is_severed) { c->p0_in = 0.02; c->p0_out = 0.00; }
    c->p_tissue = c->p0_in;
    c->p_osc = c->p0_in - c->p0_out;
    if (p.is_severed) c->p_osc *= 0.1;
    c->t_prop = (int)((c->g.nx - 1.) * sqrt(3.) - 1) * 1;
    c->t_start = 2 * c->t_prop;
    c->t_sever = 0;
    c->mp.Rho0 = 1.0 / pow(1, 3);
    const double s8 = 1.0 / p.tau, s5 = 1.0;
    const double S[9] = {1, 1, 1, 1, s5, 1, s5, s8, s8};
    memcpy(c->mp.S, S, sizeof(S));
    c->mp.alpha = p.alpha;
    c->mp.p_tissue = c->p_tissue;
}

// host-side initial state with the reference's expressions (cold path, once): Initialize_Yr_and_Vw_and_p (AB:172-189),
// Initialize_Fobj_for_Vessel_Walls (AB:275-285, mask only), Initialize_P_U_g (AB:191-214)
int initial_state(clbm_pulsatile *c, std::vector<double> &gin, std::vector<uint8_t> &flag, std::vector<double> &P,
                  std::vector<double> &Ux, std::vector<double> &Uy, std::vector<double> &yr1, std::vector<double> &yr2)
{
    const Geo &g = c->g;
    const int nx = g.nx, ny = g.ny;
    const double cc = g.c, alpha = c->prm.alpha;
    const double yr1_in = cc - (c->p0_in - c->p_tissue) / alpha, yr2_in = cc + (c->p0_in - c->p_tissue) / alpha;
    const double yr1_out = cc - (c->p0_out - c->p_tissue) / alpha, yr2_out = cc + (c->p0_out - c->p_tissue) / alpha;
    if (yr1_in < 1 || yr2_in > ny - 2 || yr1_out < 1 || yr2_out > ny - 2) {
        set_error("Initial wall location out of bounds.");
        return CLBM_EINVAL;
    }
    const double R0 = (yr2_in - yr1_in) / 2.0, RL = (yr2_out - yr1_out) / 2.0;
    auto at = [ny](int X, int Y) { return (size_t)Y + (size_t)ny * X; };
    for (int X = 0; X < nx; ++X) {
        const double Rx4 = (pow(RL, 4) - pow(R0, 4)) * ((double)X / (double)(nx - 1)) + pow(R0, 4);
        const double Rx = pow(Rx4, 0.25);
        yr1[X] = cc - Rx;
        yr2[X] = cc + Rx;
        for (int Y = 0; Y < ny; ++Y) P[at(X, Y)] = (yr2[X] - (ny - 1 - 0.5)) * alpha + c->p_tissue;
    }
    for (int X = 0; X < nx; ++X)
        for (int Y = 0; Y < ny; ++Y) flag[at(X, Y)] = fobj(yr1.data(), yr2.data(), g, X + 1, Y + 1) < 1.0 ? 0 : 1;
    const double mu = c->mp.Rho0 * (c->prm.tau - 0.5) / 3.0;
    for (int X = 0; X < nx; ++X)
        for (int Y = (int)ceil(yr1[X] - 0.01); Y <= (int)floor(yr2[X] + 0.01); ++Y) {
            double dpx;
            if (X == 0) dpx = P[at(1, Y)] - P[at(X, Y)];
            else if (X == nx - 1) dpx = P[at(X, Y)] - P[at(X - 1, Y)];
            else dpx = 0.5 * (P[at(X + 1, Y)] - P[at(X - 1, Y)]);
            Ux[at(X, Y)] = dpx / (2.0 * mu) * ((Y - yr1[X]) * (Y - yr2[X]));
        }
    for (size_t i = 0; i < (size_t)g.nelem; ++i) {
        if (flag[i] == 0) continue;   // solid: populations stay 0
        double geq[9];
        equilibrium_g(c->mp.Rho0, P[i], Ux[i], Uy[i], geq);
        for (int k = 0; k < 9; ++k) gin[(size_t)k * g.nelem + i] = geq[k];
    }
    return CLBM_OK;
}

// tensor map of one lattice buffer ([9][nx][ny], rank 3) or one field array ([nx][ny], rank 2) with a (rows, 1[, 9]) box
int puls_tmap(clbm_pulsatile *c, const void *base, int rank, int rows, CUtensorMap *out)
{
    for (const auto &e : c->tmaps)
        if (e.base == base && e.rank == rank && e.rows == rows) { memcpy(out, &e.map, sizeof(CUtensorMap)); return CLBM_OK; }
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return CLBM_ECUDA; }
    const Geo &g = c->g;
    const cuuint64_t dims[3] = {(cuuint64_t)g.ny, (cuuint64_t)g.nx, 9};
    const cuuint64_t strides[2] = {(cuuint64_t)g.ny * 8, (cuuint64_t)g.nelem * 8};
    const cuuint32_t box[3] = {(cuuint32_t)rows, 1, 9};
    const cuuint32_t estr[3] = {1, 1, 1};
    clbm_pulsatile::TmapSlot e;
    CUresult r = enc(&e.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (pulsatile, rank %d) failed (%d)", rank, (int)r); return CLBM_ECUDA; }
    e.base = base; e.rank = rank; e.rows = rows;
    c->tmaps.push_back(e);
    memcpy(out, &e.map, sizeof(CUtensorMap));
    return CLBM_OK;
}

template <int NT, int NS>
int launch_fused_tma(clbm_pulsatile *c, const double *A, double *B, double *Pw, double *Uxw, double *Uyw)
{
    using C = PulsTmaCfg<NT, NS>;
    const Geo &g = c->g;
    CUtensorMap tA, tP, tUx, tUy;
    int rc;
    if ((rc = puls_tmap(c, A, 3, C::BY, &tA)) || (rc = puls_tmap(c, c->P, 2, C::BY, &tP)) || (rc = puls_tmap(c, c->Ux, 2, C::BY, &tUx)) ||
        (rc = puls_tmap(c, c->Uy, 2, C::BY, &tUy)))
        return rc;
    auto kern = puls_fused_tma<NT, NS>;
    const int segs = (g.ny + (NT - 2) - 1) / (NT - 2);
    int xchunk = c->xchunk_env;
    if (c->tma_xchunk == 0) {       // first launch of this context: opt in to the dynamic shared memory, choose the chunk length
        CLBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        // whole waves: the CTAs of a launch all take the same time, so a last wave that is half empty costs a sixth of the launch
        int per_sm = 0, sms = 0;
        CLBM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, C::SMEM));
        CLBM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        const long long slots = (long long)(per_sm > 0 ? per_sm : 1) * sms;
        const long long want = (long long)segs * ((g.nx + 63) / 64);           // CTAs at the 64-column chunks of puls_fused
        const long long waves = (want + slots - 1) / slots;
        long long chunks = waves * slots / segs;
        if (chunks < 1) chunks = 1;
        xchunk = (int)((g.nx + chunks - 1) / chunks);
        if (xchunk < 8) xchunk = g.nx < 8 ? g.nx : 8;
        if (c->xchunk_env > 0) xchunk = c->xchunk_env;
        if (xchunk > g.nx) xchunk = g.nx;
        c->tma_xchunk = xchunk;
    }
    xchunk = c->tma_xchunk;
    dim3 grid(segs, (g.nx + xchunk - 1) / xchunk);
    kern<<<grid, NT, C::SMEM, c->stream>>>(tA, tP, tUx, tUy, B, c->flag, Pw, Uxw, Uyw, c->intr, g, c->mp, xchunk);
    return CLBM_OK;
}

int one_step(clbm_pulsatile *c)
{
    const Geo &g = c->g;
    const int t = c->t_iter;
    double *A = c->lat + (long long)c->parity * g.npop, *B = c->lat + (long long)(1 - c->parity) * g.npop;
    const int nb = grid_for(g.nelem, 256);
    const bool sample = c->ktiming && (int)c->kev.size() < 2 * c->ktiming_cap;   // event pair around the whole step
    if (sample) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->kev.push_back(e);
        cudaEventRecord(e, c->stream);
    }
    double *Pw = c->fused ? c->P2 : c->P, *Uxw = c->fused ? c->Ux2 : c->Ux, *Uyw = c->fused ? c->Uy2 : c->Uy;   // this step's P, Ux, Uy
    if (c->fused && c->tma) {
        int rc;
        if (c->tma == 3) rc = launch_fused_tma<128, 3>(c, A, B, Pw, Uxw, Uyw);
        else if (c->tma == 4) rc = launch_fused_tma<64, 3>(c, A, B, Pw, Uxw, Uyw);
        else if (c->tma == 5) rc = launch_fused_tma<254, 2>(c, A, B, Pw, Uxw, Uyw);
        else rc = launch_fused_tma<128, 2>(c, A, B, Pw, Uxw, Uyw);
        if (rc) return rc;
    } else if (c->fused) {
        int xchunk = g.nx < 64 ? g.nx : 64, nt = c->nt_env > 0 ? c->nt_env : 128;
        if (c->xchunk_env > 0) xchunk = c->xchunk_env < g.nx ? c->xchunk_env : g.nx;
        if (nt == 64) {
            dim3 grid((g.ny + 62 - 1) / 62, (g.nx + xchunk - 1) / xchunk);
            puls_fused<64><<<grid, 64, 0, c->stream>>>(A, B, c->flag, c->P, c->Ux, c->Uy, Pw, Uxw, Uyw, c->intr, g, c->mp, xchunk);
        } else {
            dim3 grid((g.ny + 126 - 1) / 126, (g.nx + xchunk - 1) / xchunk);
            puls_fused<128><<<grid, 128, 0, c->stream>>>(A, B, c->flag, c->P, c->Ux, c->Uy, Pw, Uxw, Uyw, c->intr, g, c->mp, xchunk);
        }
    } else {
        puls_collide<<<nb, 256, 0, c->stream>>>(A, B, c->flag, c->P, c->Ux, c->Uy, g, c->mp);
    }
    puls_bouzidi<<<grid_for(g.nx, 128), 128, 0, c->stream>>>(B, c->yr1, c->yr2, g, 0);
    puls_bouzidi<<<grid_for(g.nx, 128), 128, 0, c->stream>>>(B, c->yr1, c->yr2, g, 1);
    // Zou/He boundary pressures of this iteration (AB:620-622, :646-650): sin() on the host, like the reference
    double Pin = c->p0_in;
    if (t >= c->t_start) Pin = c->p0_in + c->p_osc * sin(c->omega * (t + 1 - c->t_start));
    double Pout = c->p0_out;
    if (t >= c->t_start + c->t_prop) Pout = c->p0_out + c->p_osc * sin(c->omega * (t + 1 - c->t_start - c->t_prop));
    if (t > c->t_sever) Pout = 0;
    static const int dbg = getenv("CLBM_PULS_SCAN1") ? atoi(getenv("CLBM_PULS_SCAN1")) : 0;   // 1: one node per thread in puls_stream (the first form)
    if (c->fused && !(dbg & 1))
        puls_stream<16><<<grid_for((g.nelem + 15) / 16, 256), 256, 0, c->stream>>>(c->lat, (long long)c->parity * g.npop, (long long)(1 - c->parity) * g.npop,
                                                                                 c->flag, Pw, Uxw, Uyw, c->yr1, c->yr2, g, c->mp, Pin, Pout, c->intr);
    else
        puls_stream<1><<<nb, 256, 0, c->stream>>>(c->lat, (long long)c->parity * g.npop, (long long)(1 - c->parity) * g.npop, c->flag, Pw, Uxw,
                                                  Uyw, c->yr1, c->yr2, g, c->mp, Pin, Pout, c->fused ? c->intr : nullptr);
    c->launches += 4;
    if (c->prm.deformable) {
        puls_walls<<<grid_for(g.nx, 128), 128, 0, c->stream>>>(Pw, c->yr1, c->yr2, c->yr1o, c->yr2o, g, c->mp);
        puls_fobj<<<grid_for(g.nx, 128), 128, 0, c->stream>>>(A, c->flag, Pw, Uxw, Uyw, c->yr1, c->yr2, c->yr1o, c->yr2o, g, c->mp,
                                                             c->pre, c->seed_count, c->seed_list, c->seed_cap, c->err_dev);
        puls_seed<<<1, 32, 0, c->stream>>>(A, c->flag, Pw, Uxw, Uyw, c->yr1, c->yr2, c->yr1o, c->yr2o, g, c->mp, c->pre,
                                           c->seed_count, c->seed_list, c->seed_cap, B, c->fused);
        c->launches += 3;
    }
    if (c->fused) {
        std::swap(c->P, c->P2);
        std::swap(c->Ux, c->Ux2);
        std::swap(c->Uy, c->Uy2);
        c->in_complete = false;
    }
    if (sample) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->kev.push_back(e);
        cudaEventRecord(e, c->stream);
    }
    c->parity = 1 - c->parity;
    c->t_iter++;
    CLBM_CUDA(cudaGetLastError());
    return CLBM_OK;
}

int check_device_error(clbm_pulsatile *c)
{
    CLBM_CUDA(cudaMemcpyAsync(c->err_host, c->err_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    if (*c->err_host) {
        set_error(*c->err_host == 1 ? "pulsatile: seed queue overflow (more than %d fresh nodes without an old-fluid neighbour in one step)"
                                    : "pulsatile: a wall moved by more than the +-0.25 cap allows (fresh-node window > %d rows)",
                  *c->err_host == 1 ? c->seed_cap : PRE_ROWS);
        return CLBM_ESTATE;
    }
    return CLBM_OK;
}

}  // namespace

extern "C" {

int clbm_pulsatile_create(const clbm_pulsatile_params *p, clbm_pulsatile **out)
{
    if (!p || !out) { set_error("null argument"); return CLBM_EINVAL; }
    *out = nullptr;
    if (p->abi_version != CLBM_ABI_VERSION) { set_error("ABI version %d != %d", p->abi_version, CLBM_ABI_VERSION); return CLBM_EINVAL; }
    if (p->N < 4) { set_error("N = %d too small", p->N); return CLBM_EINVAL; }
    if (!(p->tau > 0.5) || !(p->alpha > 0)) { set_error("bad tau / alpha"); return CLBM_EINVAL; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: this library has no CPU fallback");
        return CLBM_ENODEVICE;
    }
    clbm_pulsatile *c = new (std::nothrow) clbm_pulsatile();
    if (!c) { set_error("out of host memory"); return CLBM_ENOMEM; }
    c->prm = *p;
    c->g.nx = 1 + 10 * (p->N - 2);
    c->g.ny = p->N;
    c->g.Y0 = (c->g.ny - 1) / 2;
    c->g.c = c->g.Y0 + 0.5;
    c->g.nelem = (long long)c->g.nx * c->g.ny;
    c->g.npop = 9 * c->g.nelem;
    c->device = p->device;
    if (c->device < 0) cudaGetDevice(&c->device);
    if (c->device >= ndev) { set_error("device %d of %d", c->device, ndev); delete c; return CLBM_EINVAL; }
    derive_parameters(c);
    const Geo g = c->g;
    std::vector<double> gin((size_t)g.npop, 0.0), P((size_t)g.nelem, 0.0), Ux((size_t)g.nelem, 0.0), Uy((size_t)g.nelem, 0.0), yr1(g.nx), yr2(g.nx);
    std::vector<uint8_t> flag((size_t)g.nelem, 1);
    int rc = initial_state(c, gin, flag, P, Ux, Uy, yr1, yr2);
    if (rc) { delete c; return rc; }
    auto fail = [&](cudaError_t e, const char *what) { int r = cuda_fail(e, what, __FILE__, __LINE__); clbm_pulsatile_destroy(c); return r; };
    cudaError_t e;
    if ((e = cudaSetDevice(c->device)) != cudaSuccess) return fail(e, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    const size_t nd = (size_t)g.nelem * sizeof(double);
    if ((e = cudaMalloc(&c->lat, (2 * (size_t)g.npop + 1) * sizeof(double))) != cudaSuccess) return fail(e, "cudaMalloc lattice");
    if ((e = cudaMalloc(&c->flag, (size_t)g.nelem)) != cudaSuccess) return fail(e, "cudaMalloc flag");
    if ((e = cudaMalloc(&c->intr, (size_t)g.nelem)) != cudaSuccess) return fail(e, "cudaMalloc intr");
    c->fused = 1;
    if (const char *e = getenv("CLBM_PULS_FUSED")) c->fused = atoi(e) != 0;
    c->tma_xchunk = 0;
    c->xchunk_env = getenv("CLBM_PULS_XCHUNK") ? atoi(getenv("CLBM_PULS_XCHUNK")) : 0;
    c->nt_env = getenv("CLBM_PULS_NT") ? atoi(getenv("CLBM_PULS_NT")) : 0;
    // TMA-staged fused step, OPT-IN (measured equal or slower, see puls_fused_tma): CLBM_PULS_TMA = 2 / 3 pick the stage count of the
    // 128-row tile, 4 the 64-row tile, 5 the 254-row tile; needs 16-byte aligned rows and buffers (even N) and the driver's encoder
    c->tma = 0;
    if (const char *e = getenv("CLBM_PULS_TMA")) {
        const int v = atoi(e);
        if (v >= 2 && v <= 5 && g.ny % 2 == 0 && g.ny >= 8 && get_encode() != nullptr) c->tma = v;
    }
    c->in_complete = true;
    double **fl[6] = {&c->P, &c->Ux, &c->Uy, &c->P2, &c->Ux2, &c->Uy2};
    for (auto f : fl) if ((e = cudaMalloc(f, nd)) != cudaSuccess) return fail(e, "cudaMalloc field");
    double **wl[4] = {&c->yr1, &c->yr2, &c->yr1o, &c->yr2o};
    for (auto w : wl) if ((e = cudaMalloc(w, g.nx * sizeof(double))) != cudaSuccess) return fail(e, "cudaMalloc wall");
    if ((e = cudaMalloc(&c->err_dev, sizeof(int))) != cudaSuccess) return fail(e, "cudaMalloc err");
    c->seed_cap = 4 * g.nx + 64;
    if ((e = cudaMalloc(&c->pre, (size_t)g.nx * 2 * PRE_ROWS * 9 * sizeof(double))) != cudaSuccess) return fail(e, "cudaMalloc pre");
    if ((e = cudaMalloc(&c->seed_count, sizeof(int))) != cudaSuccess) return fail(e, "cudaMalloc seed_count");
    if ((e = cudaMalloc(&c->seed_list, c->seed_cap * sizeof(int))) != cudaSuccess) return fail(e, "cudaMalloc seed_list");
    cudaMemsetAsync(c->seed_count, 0, sizeof(int), c->stream);
    if ((e = cudaMallocHost(&c->err_host, sizeof(int))) != cudaSuccess) return fail(e, "cudaMallocHost err");
    cudaMemsetAsync(c->lat, 0, (2 * (size_t)g.npop + 1) * sizeof(double), c->stream);
    cudaMemsetAsync(c->err_dev, 0, sizeof(int), c->stream);
    cudaMemcpyAsync(c->lat, gin.data(), (size_t)g.npop * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(c->flag, flag.data(), (size_t)g.nelem, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(c->P, P.data(), nd, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(c->Ux, Ux.data(), nd, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(c->Uy, Uy.data(), nd, cudaMemcpyHostToDevice, c->stream);
    for (int w = 0; w < 2; ++w) {
        cudaMemcpyAsync(w ? c->yr1o : c->yr1, yr1.data(), g.nx * sizeof(double), cudaMemcpyHostToDevice, c->stream);
        cudaMemcpyAsync(w ? c->yr2o : c->yr2, yr2.data(), g.nx * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    }
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return fail(e, "initial upload");
    *out = c;
    return CLBM_OK;
}

int clbm_pulsatile_destroy(clbm_pulsatile *c)
{
    if (!c) return CLBM_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->lat); cudaFree(c->flag); cudaFree(c->P); cudaFree(c->Ux); cudaFree(c->Uy);
    cudaFree(c->P2); cudaFree(c->Ux2); cudaFree(c->Uy2); cudaFree(c->intr);
    cudaFree(c->yr1); cudaFree(c->yr2); cudaFree(c->yr1o); cudaFree(c->yr2o); cudaFree(c->err_dev);
    cudaFree(c->pre); cudaFree(c->seed_count); cudaFree(c->seed_list);
    if (c->err_host) cudaFreeHost(c->err_host);
    for (auto e : c->kev) cudaEventDestroy(e);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return CLBM_OK;
}

int clbm_pulsatile_info(const clbm_pulsatile *c, int *nx, int *ny, int *tf, int *t_iter, int *parity)
{
    if (!c) { set_error("null context"); return CLBM_EINVAL; }
    if (nx) *nx = c->g.nx;
    if (ny) *ny = c->g.ny;
    if (tf) *tf = c->t_beat + 2 * c->t_prop;     // AB:759
    if (t_iter) *t_iter = c->t_iter;
    if (parity) *parity = c->parity;
    return CLBM_OK;
}

int clbm_pulsatile_step(clbm_pulsatile *c, int nsteps)
{
    if (!c || nsteps < 0) { set_error("bad argument"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    for (int s = 0; s < nsteps; ++s) {
        int rc = one_step(c);
        if (rc) return rc;
    }
    return CLBM_OK;
}

int clbm_pulsatile_sync(clbm_pulsatile *c)
{
    if (!c) { set_error("null context"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    return check_device_error(c);
}

int clbm_pulsatile_step_timed(clbm_pulsatile *c, int nsteps, float *ms)
{
    if (!c || nsteps < 0 || !ms) { set_error("bad argument"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    CLBM_CUDA(cudaEventRecord(c->ev0, c->stream));
    for (int s = 0; s < nsteps; ++s) {
        int rc = one_step(c);
        if (rc) return rc;
    }
    CLBM_CUDA(cudaEventRecord(c->ev1, c->stream));
    CLBM_CUDA(cudaEventSynchronize(c->ev1));
    CLBM_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return check_device_error(c);
}

int64_t clbm_pulsatile_launch_count(const clbm_pulsatile *c) { return c ? c->launches : 0; }

int clbm_pulsatile_kernel_timing_begin(clbm_pulsatile *c, int cap_steps)
{
    if (!c) { set_error("null context"); return CLBM_EINVAL; }
    for (auto e : c->kev) cudaEventDestroy(e);
    c->kev.clear();
    c->ktiming = true;
    c->ktiming_cap = cap_steps;
    return CLBM_OK;
}

int clbm_pulsatile_kernel_timing_end(clbm_pulsatile *c, float *avg_ms, int *count)
{
    if (!c || !avg_ms || !count) { set_error("bad argument"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    double sum = 0;
    int n = 0;
    for (size_t i = 0; i + 1 < c->kev.size(); i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->kev[i], c->kev[i + 1]) == cudaSuccess) { sum += ms; ++n; }
    }
    for (auto e : c->kev) cudaEventDestroy(e);
    c->kev.clear();
    c->ktiming = false;
    *avg_ms = n ? (float)(sum / n) : 0.f;
    *count = n;
    return CLBM_OK;
}

int clbm_pulsatile_download_fields(clbm_pulsatile *c, double *P, double *Ux, double *Uy, uint8_t *flag, double *yr1, double *yr2)
{
    if (!c) { set_error("null context"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    const size_t nd = (size_t)c->g.nelem * sizeof(double);
    if (P) CLBM_CUDA(cudaMemcpyAsync(P, c->P, nd, cudaMemcpyDeviceToHost, c->stream));
    if (Ux) CLBM_CUDA(cudaMemcpyAsync(Ux, c->Ux, nd, cudaMemcpyDeviceToHost, c->stream));
    if (Uy) CLBM_CUDA(cudaMemcpyAsync(Uy, c->Uy, nd, cudaMemcpyDeviceToHost, c->stream));
    if (flag) CLBM_CUDA(cudaMemcpyAsync(flag, c->flag, (size_t)c->g.nelem, cudaMemcpyDeviceToHost, c->stream));
    if (yr1) CLBM_CUDA(cudaMemcpyAsync(yr1, c->yr1, c->g.nx * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (yr2) CLBM_CUDA(cudaMemcpyAsync(yr2, c->yr2, c->g.nx * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return check_device_error(c);
}

int clbm_pulsatile_download_lattice(clbm_pulsatile *c, double *lattice, int *parity)
{
    if (!c || !lattice) { set_error("bad argument"); return CLBM_EINVAL; }
    CLBM_CUDA(cudaSetDevice(c->device));
    if (!c->in_complete) {
        // the last step's in buffer is the one the parity flip turned into the out buffer
        double *A = c->lat + (long long)(1 - c->parity) * c->g.npop, *B = c->lat + (long long)c->parity * c->g.npop;
        puls_materialise<<<grid_for(c->g.nelem, 256), 256, 0, c->stream>>>(A, B, c->yr1o, c->yr2o, c->g);
        CLBM_CUDA(cudaGetLastError());
        c->launches += 1;
        c->in_complete = true;
    }
    CLBM_CUDA(cudaMemcpyAsync(lattice, c->lat, 2 * (size_t)c->g.npop * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (parity) *parity = c->parity;
    return check_device_error(c);
}

int clbm_pulsatile_upload(clbm_pulsatile *c, const double *lattice, const uint8_t *flag, const double *P, const double *Ux,
                          const double *Uy, const double *yr1, const double *yr2, int parity, int t_iter)
{
    if (!c || !lattice || !flag || !P || !Ux || !Uy || !yr1 || !yr2 || (parity != 0 && parity != 1) || t_iter < 0) {
        set_error("bad argument");
        return CLBM_EINVAL;
    }
    CLBM_CUDA(cudaSetDevice(c->device));
    const size_t nd = (size_t)c->g.nelem * sizeof(double);
    CLBM_CUDA(cudaMemcpyAsync(c->lat, lattice, 2 * (size_t)c->g.npop * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CLBM_CUDA(cudaMemcpyAsync(c->flag, flag, (size_t)c->g.nelem, cudaMemcpyHostToDevice, c->stream));
    CLBM_CUDA(cudaMemcpyAsync(c->P, P, nd, cudaMemcpyHostToDevice, c->stream));
    CLBM_CUDA(cudaMemcpyAsync(c->Ux, Ux, nd, cudaMemcpyHostToDevice, c->stream));
    CLBM_CUDA(cudaMemcpyAsync(c->Uy, Uy, nd, cudaMemcpyHostToDevice, c->stream));
    CLBM_CUDA(cudaMemcpyAsync(c->yr1, yr1, c->g.nx * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CLBM_CUDA(cudaMemcpyAsync(c->yr2, yr2, c->g.nx * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CLBM_CUDA(cudaMemcpyAsync(c->yr1o, yr1, c->g.nx * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CLBM_CUDA(cudaMemcpyAsync(c->yr2o, yr2, c->g.nx * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CLBM_CUDA(cudaStreamSynchronize(c->stream));
    c->in_complete = true;
    c->parity = parity;
    c->t_iter = t_iter;
    return CLBM_OK;
}

}  // extern "C"
