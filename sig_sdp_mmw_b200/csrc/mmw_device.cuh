// Device side of the MMW hot path: every phase of one iteration as a __device__
// function over a "team" of thread blocks, so the same code runs
//   * fused    : one persistent cooperative kernel, grid-wide barriers between phases,
//   * stepwise : one kernel per phase (kernel boundaries are the barriers),
//   * batch    : one thread block per independent instance (__syncthreads barriers).
//
// Arithmetic follows SURVEY.md App. A (reference sim_src/alg/mmw.py:77-197 and scipy's
// expm_multiply, _expm_multiply.py:214-303); each phase cites its lines.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cg = cooperative_groups;

namespace sigsdp {

constexpr int NT = 512;          // threads per block
constexpr int NWARP = NT / 32;
constexpr int HIST = 8192;       // Taylor-controller history ring
constexpr int PSTRIDE = 8;       // doubles per block in the partial-sum scratch

// ---------------------------------------------------------------------------
struct PlanDev {
    int n, nnz, E_g, E_a;
    const int* rowptr;      // n + 1
    const int* col;         // nnz, ascending inside a row, diagonal included
    const int* eid;         // nnz: -1 diag, [0,E_g) gain edge, E_g + a asso edge
    const double* tfwd;     // nnz: T[row,col]
    const double* tbwd;     // nnz: T[col,row]
    const double* S_sum;    // n
    const double* tnorm;    // n  sqrt(sum_c T[k,c]^2)
    const double* h_max;    // n
    const int* perm;        // n: internal -> caller numbering (nullptr = identity)
    const int* dpos;        // n: position of the diagonal entry of each row
    const int* apos;        // E_a: position of the (row < col) entry of each asso edge
};

// Row tiles for the staged kernels: the rows of a tile are consecutive (after the
// locality renumbering they are a compact patch of the graph), `ucol` lists the distinct
// columns the tile touches (ascending) and `lcol` maps every non-zero to its index in
// that list, so a tile's slice of the sketch block can be staged in shared memory once
// and reused by all of its rows.
struct TileDev {
    int enabled;                 // 0 = no tiling (direct-gather kernels)
    int ntiles, ucap, nnzcap;
    const int* trow;             // ntiles + 1: first row of each tile
    const int* ucnt;             // ntiles: distinct columns of the tile
    const int* rptr;             // ntiles + 1: runs of consecutive distinct columns
    const int4* runs;            // {first column, first shared-memory slot, length, 0}
    const unsigned short* lcol;  // nnz (+ padding)
    const int4* trec;            // 2 per tile: {r0, r1, p0, p1}, {run0, run1, distinct columns, 0}
    // phase_term_staged2 / phase_gram_staged2 only (per solver, nullptr otherwise): what each of the
    // slot_r lane groups of a block does in a tile, {row or -1, first non-zero, non-zeros | role << 16,
    // position of the row's diagonal entry}.  A tile has
    // fewer rows than groups; the spare groups take the second halves of its longest rows
    // (role 1 = first half, adds the partial sums of the group to its right; 2 = second half, no
    // epilogue; 0 = whole row), so the multiply ends with the longest HALF row.
    const int4* slots;
    int slot_r;
};

// Row sharding of ONE graph across the GPUs of a box (BASELINE configs[3]): rank r owns the
// consecutive rows [row_lo, row_hi) (whole tiles) of the locality-ordered pattern -- their
// L_accu / X entries, dual weights and sketch rows.  Every rank keeps full-length arrays in the
// plan's global numbering and only walks its own part; the rows of B / F (per Taylor term),
// r and q (per iteration) that a neighbouring rank's rows read are pushed into that rank's
// copy of the array by the kernel that produces them, with plain stores through peer-mapped
// memory (NVLink), and become visible at the next team barrier.  nranks == 1: everything is
// "own" and none of this is touched.
constexpr int MAXR = 8;          // ranks of a row-sharded solver
struct ShardDev {
    int nranks, rank;
    int row_lo, row_hi;          // own rows
    int tile_lo, tile_hi;        // own tiles of the staged kernels
    // association edges with an entry in an own row: the first n_inc_owned are owned (their
    // row < col entry is in an own row: this rank accounts for them in the soft-max sums), the
    // rest belong to a neighbour and are kept up to date redundantly (L needs their weight)
    int n_inc, n_inc_owned;
    const int* inc_e;            // asso edge id (nullptr: edge i, position apos[i])
    const int* inc_pos;          // position of that edge's entry in an own row
    const unsigned char* pmask;  // n: bit p set = rank p reads row k of B / F / r / q (nullptr: unsharded)
    long long delta[MAXR];       // bytes from this rank's exchange arena to rank p's
    unsigned long long* flags;   // (reserved: first 256 bytes of the arena)
    unsigned long long* inbox;   // [2][MAXR][16] the ranks' scalar words {epoch | half}, by barrier parity
    unsigned long long timeout_ns;
};

// device-resident controller: reductions that must be order independent use
// atomicMax on the bit pattern of non-negative doubles
struct Ctrl {
    unsigned long long nrm_b[3];   // ||B_new||_inf per term slot
    unsigned long long nrm_f[3];   // ||F||_inf per term slot
    unsigned long long a1_key;     // ||A - mu I||_1
    unsigned long long c1_key;     // ||Omega_hat||_inf
    unsigned long long emax_key;   // max e_accu (order-preserving key)
    double trL[2];                 // tr(L_accu), double-buffered by iteration parity
    double mu;                     // shift of the current sketch (for exporting Y_h)
    double smax_shift;             // soft-max shift of the next iteration: max e_accu of the last one (0 after reset)
    long long total_terms;
    long long iter;                // iterations done since reset
    // stepwise-mode controller (host reads these back)
    int m_star, done, nterms;
    long long s;
    double c1, a1;
    // diagnostics: [4] nanoseconds the leader thread spent in team barriers (the others are unused)
    long long dbg[8];
    unsigned long long xepoch;     // cross-GPU barriers completed so far (row-sharded solvers)
    // arrivals at the fused kernel's grid barrier, counted up for ever (GridTeam); on a line
    // of its own so the polling does not collide with the keys above
    alignas(128) unsigned long long bar;
    unsigned long long bar_pad[15];
};

template <typename T>
struct Prob {
    PlanDev g;
    int Z, D, Dp, C;
    double eta;
    double tol;
    const double* nH;      // n  norm_H for this Z (mmw.py:39)
    const double* hcoef;   // n  h/K - S_sum/(K Z)
    double* Lval;          // nnz  L_accu on the union pattern
    T* Aval;               // nnz  L_accu/2 - mu I in the sketch dtype (what the Taylor terms multiply by)
    double* e_acc;         // C
    double* u;             // C    exp(e_acc - max), unnormalised
    double* Y;             // C
    double* Ybar;          // C    running sum
    double* q;             // n    u_H / norm_H
    double* Xv;            // nnz  X on the union pattern (every directed entry, diagonal included)
    double* Xbarv;         // nnz  running sum of X
    double* r;             // n    row sums of off-diagonal X
    double* dsq;           // n    ||F_k||^2
    T* B0;                 // n x Dp ping
    T* B1;                 // n x Dp pong
    T* F;                  // n x Dp
    double* psum;          // [blocks][PSTRIDE] softmax sums
    double* ptr;           // [blocks] trace partials
    Ctrl* ctrl;
    int* hist_m;           // HIST
    int* hist_s;
    int* hist_nt;
    double* hist_a1;
    double* hist_mu;
    double* hist_t;        // HIST x 4: device-timed microseconds of dual / loss / Taylor terms / Gram
    const double* omega;   // raw normals for this call (caller numbering) or nullptr
    unsigned long long seed;
    TileDev tl;
    // sketch-column sharding across GPUs: this solver holds columns [col0, col0 + D) of a
    // Dtot-wide sketch.  split != 0: the Gram phase writes un-normalised partial dot products
    // to graw (the terms' partial ||F_k||^2 go to dsq, which then lives right behind graw);
    // the ranks all-reduce that buffer and phase_gram_finish completes X, X_avgd and r.
    int Dtot, col0, split;
    double* graw;          // nnz (+ n: dsq)
    ShardDev sh;           // row sharding (nranks == 1: none)
};

// ---------------------------------------------------------------------------
// small helpers
__device__ __forceinline__ unsigned long long dkey_pos(double v) {  // v >= 0
    return (unsigned long long)__double_as_longlong(v);
}
__device__ __forceinline__ double dkey_pos_inv(unsigned long long k) { return __longlong_as_double((long long)k); }
__device__ __forceinline__ unsigned long long dkey_any(double v) {  // total order on all doubles
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_any_inv(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long ld_u64(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ double ld_f64(const double* p) { return *reinterpret_cast<const volatile double*>(p); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// deterministic block reductions; result valid in every thread
__device__ __forceinline__ double block_sum(double v, double* sh /* NWARP + 1 */) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < NWARP ? sh[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) sh[NWARP] = t;
    }
    __syncthreads();
    return sh[NWARP];
}
__device__ __forceinline__ double block_max(double v, double* sh) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < NWARP ? sh[threadIdx.x] : -INFINITY;
        t = warp_max(t);
        if (threadIdx.x == 0) sh[NWARP] = t;
    }
    __syncthreads();
    return sh[NWARP];
}
// One warp sums part[b * stride], b < nblk, in a fixed order with eight independent loads in flight
// per lane (a dynamic-trip-count loop of dependent loads would expose one L2 latency per 32 blocks:
// measured 5 us for 296 blocks).  Result valid in every lane.
__device__ __forceinline__ double warp_strided_sum(const double* part, int stride, int nblk, int lane) {
    double t = 0.0;
    for (int b0 = 0; b0 < nblk; b0 += 256) {
        double v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int b = b0 + 32 * i + lane;
            v[i] = b < nblk ? ld_f64(part + (size_t)b * stride) : 0.0;
        }
        t += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
    return warp_sum(t);
}
// sum of one double per block written by every block of the team before the last
// barrier; fixed order, so every block (and every run) gets the same bits
__device__ __forceinline__ double team_sum(const double* part, int stride, int nblk, double* sh) {
    if (threadIdx.x < 32) {
        const double t = warp_strided_sum(part, stride, nblk, threadIdx.x);
        if (threadIdx.x == 0) sh[NWARP] = t;
    }
    __syncthreads();
    const double t = sh[NWARP];
    __syncthreads();
    return t;
}
// four such sums (columns 0..3 of a [blocks][PSTRIDE] array) by four warps at once
__device__ __forceinline__ void team_sum4(const double* part, int nblk, double* sh, double s[4]) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w < 4) {
        const double t = warp_strided_sum(part + w, PSTRIDE, nblk, lane);
        if (lane == 0) sh[w] = t;
    }
    __syncthreads();
    s[0] = sh[0];
    s[1] = sh[1];
    s[2] = sh[2];
    s[3] = sh[3];
    __syncthreads();
}

template <int G>
__device__ __forceinline__ double group_sum(cg::thread_block_tile<G>& tile, double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += tile.shfl_xor(v, o);
    return v;
}

// ---------------------------------------------------------------------------
// teams: the set of thread blocks that runs one solver, its barrier, and how a phase reads
// the scalars the previous phase reduced (max e_accu, soft-max sums, ||A||_1, ||.||_inf of
// the Taylor terms, trace).  The phases are written against this interface:
//   sync(P, what, slot)   barrier; `what` names the scalars the finished phase produced
//   emax / exp_sums / a1 / c1 / trace_sum / term_norms   those scalars, valid after the sync
enum SyncWhat { SY_PLAIN = 0, SY_DUAL = 1, SY_LOSS = 3, SY_TERM = 4 };

// spin-wait helper: tight polls first, then back off; a barrier that does not complete
// within `limit_ns` (a block died, a peer never launched) traps instead of hanging the GPU
struct SpinGuard {
    unsigned n = 0;
    unsigned long long t0 = 0;
    __device__ __forceinline__ void step(unsigned long long limit_ns) {
        if (++n < 2048u) return;
        if (t0 == 0) t0 = globaltimer_ns();
        __nanosleep(64);
        if ((n & 1023u) == 0u && limit_ns && globaltimer_ns() - t0 > limit_ns) __trap();
    }
};
#ifndef SIGSDP_LOCAL_BARRIER_TIMEOUT_NS
#define SIGSDP_LOCAL_BARRIER_TIMEOUT_NS 20000000000ull   // 20 s: far beyond any phase; 0 disables
#endif

// shared by the single-GPU teams: the scalars live in this solver's Ctrl / partial arrays
struct LocalScalars {
    __device__ void note_push() const {}
    template <typename T> __device__ double emax(const Prob<T>& P) const { return dkey_any_inv(ld_u64(&P.ctrl->emax_key)); }
    template <typename T> __device__ void exp_sums(const Prob<T>& P, int nblk, double* sh, double s[4]) const {
        team_sum4(P.psum, nblk, sh, s);
    }
    template <typename T> __device__ double a1(const Prob<T>& P) const { return dkey_pos_inv(ld_u64(&P.ctrl->a1_key)); }
    template <typename T> __device__ double c1(const Prob<T>& P) const { return dkey_pos_inv(ld_u64(&P.ctrl->c1_key)); }
    template <typename T> __device__ void term_norms(const Prob<T>& P, int slot, double& c2, double& fn) const {
        c2 = dkey_pos_inv(ld_u64(&P.ctrl->nrm_b[slot]));
        fn = dkey_pos_inv(ld_u64(&P.ctrl->nrm_f[slot]));
    }
};

// The fused kernel's grid barrier (cooperative launch: all blocks are resident).  One
// arrival counter that only counts up: barrier number b is complete when it reaches
// b * gridDim.x, so there is no reset and no second round trip; every block derives the
// number it starts at from the counter itself (arrivals of the first barrier of a launch
// cannot complete it before this block has arrived too, so rounding down is exact).
// Thread 0 arrives with a release and polls with an acquire at GPU scope; the block
// barriers on either side extend that to the whole block.
struct GridTeam : LocalScalars {
    unsigned long long* ctr;
    mutable unsigned long long next;
    __device__ explicit GridTeam(unsigned long long* c) : ctr(c) {
        unsigned long long v;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(c) : "memory");
        next = (v / gridDim.x + 1ull) * gridDim.x;
    }
    __device__ int rank() const { return blockIdx.x; }
    __device__ int size() const { return gridDim.x; }
    template <typename T> __device__ void sync(const Prob<T>&, int = SY_PLAIN, int = 0) const {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(ctr) : "memory");
            unsigned long long v;
            SpinGuard g;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
                if (v >= next) break;
                g.step(SIGSDP_LOCAL_BARRIER_TIMEOUT_NS);
            }
        }
        next += gridDim.x;
        __syncthreads();
    }
    template <typename T> __device__ double trace_sum(const Prob<T>& P, double* sh) const { return team_sum(P.ptr, 1, size(), sh); }
    template <typename T> __device__ void exp_sums(const Prob<T>& P, double* sh, double s[4]) const { LocalScalars::exp_sums(P, size(), sh, s); }
    template <typename T> __device__ void finish(const Prob<T>&) const {}
};
struct StepTeam : LocalScalars {  // stepwise: the kernel boundary is the barrier
    __device__ int rank() const { return blockIdx.x; }
    __device__ int size() const { return gridDim.x; }
    template <typename T> __device__ void sync(const Prob<T>&, int = SY_PLAIN, int = 0) const {}
    template <typename T> __device__ double trace_sum(const Prob<T>& P, double* sh) const { return team_sum(P.ptr, 1, size(), sh); }
    template <typename T> __device__ void exp_sums(const Prob<T>& P, double* sh, double s[4]) const { LocalScalars::exp_sums(P, size(), sh, s); }
    template <typename T> __device__ void finish(const Prob<T>&) const {}
};
struct CtaTeam : LocalScalars {  // batch: one block owns the instance
    __device__ int rank() const { return 0; }
    __device__ int size() const { return 1; }
    template <typename T> __device__ void sync(const Prob<T>&, int = SY_PLAIN, int = 0) const { __syncthreads(); }
    template <typename T> __device__ double trace_sum(const Prob<T>& P, double* sh) const { return team_sum(P.ptr, 1, 1, sh); }
    template <typename T> __device__ void exp_sums(const Prob<T>& P, double* sh, double s[4]) const { LocalScalars::exp_sums(P, 1, sh, s); }
    template <typename T> __device__ void finish(const Prob<T>&) const {}
};

// Batch with several blocks per instance: the B consecutive blocks of one instance form a small
// team with the fused kernel's counting barrier on that instance's own counter (cooperative
// launch: all blocks resident).  Used when a batch has fewer instances than the GPU has block
// slots (128 instances per GPU on an 8-GPU split leave 57 % of a B200 idle with one block each).
struct BatchTeam : LocalScalars {
    unsigned long long* ctr;
    mutable unsigned long long next;
    int B, r;
    __device__ BatchTeam(unsigned long long* c, int B_, int r_) : ctr(c), B(B_), r(r_) {
        unsigned long long v;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(c) : "memory");
        next = (v / (unsigned)B + 1ull) * (unsigned)B;
    }
    __device__ int rank() const { return r; }
    __device__ int size() const { return B; }
    template <typename T> __device__ void sync(const Prob<T>&, int = SY_PLAIN, int = 0) const {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(ctr) : "memory");
            unsigned long long v;
            SpinGuard g;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
                if (v >= next) break;
                g.step(SIGSDP_LOCAL_BARRIER_TIMEOUT_NS);
            }
        }
        next += (unsigned)B;
        __syncthreads();
    }
    template <typename T> __device__ double trace_sum(const Prob<T>& P, double* sh) const { return team_sum(P.ptr, 1, B, sh); }
    template <typename T> __device__ void exp_sums(const Prob<T>& P, double* sh, double s[4]) const { LocalScalars::exp_sums(P, B, sh, s); }
    template <typename T> __device__ void finish(const Prob<T>&) const {}
};

// Row-sharded solver: the blocks of this GPU plus, through peer-mapped memory, the blocks of
// the other ranks.  A barrier is
//  (1) every block of this GPU arrives on the local counter; a block that pushed halo rows into
//      a peer's arrays since the last barrier first orders them with a system-scope fence (the
//      stores have reached the peer), the others only need a GPU-scope fence;
//  (2) block 0 waits for the local count, reduces this rank's partial scalars in a fixed order
//      and stores them into slot [parity][rank] of EVERY rank's inbox as 16 self-validating
//      words {epoch : 32 | half of a scalar : 32} (the "LL" protocol of collective libraries: an
//      8-byte store is atomic, so no fence is needed between payload and flag and the words
//      double as the arrival flags);
//  (3) every block of every rank polls its own inbox until all ranks' words carry the epoch,
//      then reduces the scalars over the ranks in rank order -- so every block on every GPU
//      continues with bit-identical scalars (same Taylor degree, same early exit, same trace),
//      which is what keeps the ranks' control flow in step without a host.
// The inbox is double-buffered by barrier parity: a rank can be at most one barrier ahead.
// Measured on 2 x B200 (scripts/micro/p2p_fence.cu): fence.sys 0.9 us idle, 1.7 us after a peer
// store, one-way peer store 1.2 us, red.release.sys 1.8 us vs red.release.gpu 0.4 us.
constexpr int LLW = 16;   // LL words per rank and barrier: 8 scalars x 2 halves
struct ShardTeam {
    unsigned long long* ctr;
    mutable unsigned long long next;
    mutable unsigned long long epoch;
    mutable int pushed;
    double* xr;      // shared memory, 8 doubles: [0..3] keys (as doubles' bit patterns), [4..7] sums
    unsigned* xh;    // shared memory, MAXR * LLW halves as received
    template <typename T>
    __device__ ShardTeam(const Prob<T>& P, double* xr_, unsigned* xh_) : ctr(&P.ctrl->bar), pushed(0), xr(xr_), xh(xh_) {
        unsigned long long v;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
        next = (v / gridDim.x + 1ull) * gridDim.x;
        epoch = *reinterpret_cast<const volatile unsigned long long*>(&P.ctrl->xepoch);
    }
    __device__ int rank() const { return blockIdx.x; }
    __device__ int size() const { return gridDim.x; }
    __device__ void note_push() const { pushed = 1; }
    template <typename T>
    __device__ void sync(const Prob<T>& P, int what = SY_PLAIN, int slot = 0) const {
        const ShardDev& S = P.sh;
        Ctrl* ctrl = P.ctrl;
        const unsigned long long ep = epoch + 1ull;
        const unsigned par = (unsigned)(ep & 1ull);
        const unsigned ep32 = (unsigned)ep;
        const int any_push = __syncthreads_or(pushed);
        pushed = 0;
        if (threadIdx.x == 0) {
            if (any_push) __threadfence_system(); else __threadfence();
            asm volatile("red.relaxed.gpu.global.add.u64 [%0], 1;" ::"l"(ctr) : "memory");
        }
        unsigned long long tb0 = 0, tb1 = 0, tb2 = 0;   // leader's timestamps: arrived / local complete / words sent
        if (blockIdx.x == 0 && threadIdx.x < 32) {
            const int lane = threadIdx.x;
            if (lane == 0) {
                tb0 = globaltimer_ns();
                unsigned long long v;
                SpinGuard g;
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
                    if (v >= next) break;
                    g.step(S.timeout_ns);
                }
                tb1 = globaltimer_ns();
            }
            __syncwarp();
            // this rank's payload: 4 keys (max-reduced across ranks) and 4 sums
            unsigned long long pk[4] = {0ull, 0ull, 0ull, 0ull};
            double ps[4] = {0.0, 0.0, 0.0, 0.0};
            const int nblk = gridDim.x;
            if (what == SY_DUAL) {
                pk[0] = ld_u64(&ctrl->emax_key);
#pragma unroll
                for (int i = 0; i < 4; ++i) ps[i] = warp_strided_sum(P.psum + i, PSTRIDE, nblk, lane);
            } else if (what == SY_LOSS || what == SY_TERM) {
                pk[0] = what == SY_LOSS ? ld_u64(&ctrl->a1_key) : ld_u64(&ctrl->nrm_b[slot]);
                pk[1] = what == SY_LOSS ? ld_u64(&ctrl->c1_key) : ld_u64(&ctrl->nrm_f[slot]);
                ps[0] = warp_strided_sum(P.ptr, 1, nblk, lane);
            }
            // lane l < 16 sends half (l & 1) of scalar l >> 1, tagged with the epoch
            const int si = lane >> 1;
            unsigned long long full = 0ull;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (si == i) full = pk[i];
                if (si == 4 + i) full = (unsigned long long)__double_as_longlong(ps[i]);
            }
            const unsigned half = (lane & 1) ? (unsigned)(full >> 32) : (unsigned)full;
            const unsigned long long word = ((unsigned long long)ep32 << 32) | (unsigned long long)half;
            // release: what this GPU's blocks published before arriving is ordered before the words
            // (halo rows were already fenced at system scope by the blocks that wrote them)
            __threadfence_system();
            if (lane < LLW) {
                for (int r = 0; r < S.nranks; ++r) {
                    unsigned long long* dst = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(S.inbox) + S.delta[r]) +
                                              ((size_t)par * MAXR + S.rank) * LLW + lane;
                    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
                }
            }
            __syncwarp();
            if (lane == 0) tb2 = globaltimer_ns();
        }
        if (threadIdx.x < LLW * S.nranks) {   // one reader per word; the acquire load also drops stale L1 lines
            // only the reader of a rank's LAST word polls (with a short sleep: thousands of tight
            // pollers slow the leader's own memory operations down); when that word is there the
            // other fifteen almost always are, and every reader still checks its own word
            const unsigned grp_mask = __activemask();
            const unsigned long long* src = S.inbox + ((size_t)par * MAXR + (threadIdx.x / LLW)) * LLW + (threadIdx.x % LLW);
            unsigned long long v;
            SpinGuard g;
            if ((threadIdx.x % LLW) == LLW - 1) {
                for (;;) {
                    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
                    if ((unsigned)(v >> 32) == ep32) break;
                    __nanosleep(40);
                    g.step(S.timeout_ns);
                }
            }
            __syncwarp(grp_mask);
            for (;;) {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
                if ((unsigned)(v >> 32) == ep32) break;
                g.step(S.timeout_ns);
            }
            xh[threadIdx.x] = (unsigned)v;
        }
        __syncthreads();
        if (what != SY_PLAIN && threadIdx.x < 8) {
            if (threadIdx.x < 4) {
                unsigned long long m = 0ull;
                for (int r = 0; r < S.nranks; ++r) {
                    const unsigned long long v = ((unsigned long long)xh[r * LLW + 2 * threadIdx.x + 1] << 32) | xh[r * LLW + 2 * threadIdx.x];
                    m = v > m ? v : m;
                }
                xr[threadIdx.x] = __longlong_as_double((long long)m);
            } else {
                double a = 0.0;
                for (int r = 0; r < S.nranks; ++r) {
                    const unsigned long long v = ((unsigned long long)xh[r * LLW + 2 * threadIdx.x + 1] << 32) | xh[r * LLW + 2 * threadIdx.x];
                    a += __longlong_as_double((long long)v);
                }
                xr[threadIdx.x] = a;
            }
        }
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0) {   // where the leader's barrier time goes (ns, see sigsdp_solver_debug_cycles)
            const unsigned long long tb3 = globaltimer_ns();
            ctrl->dbg[5] += (long long)(tb1 - tb0);   // waiting for this GPU's blocks
            ctrl->dbg[6] += (long long)(tb2 - tb1);   // reducing + sending the scalar words to the peers
            ctrl->dbg[7] += (long long)(tb3 - tb2);   // waiting for the peers' words
        }
        next += gridDim.x;
        epoch = ep;
    }
    __device__ unsigned long long key(int i) const { return (unsigned long long)__double_as_longlong(xr[i]); }
    template <typename T> __device__ double emax(const Prob<T>&) const { return dkey_any_inv(key(0)); }
    template <typename T> __device__ void exp_sums(const Prob<T>&, double*, double s[4]) const {
        s[0] = xr[4]; s[1] = xr[5]; s[2] = xr[6]; s[3] = xr[7];
    }
    template <typename T> __device__ double a1(const Prob<T>&) const { return dkey_pos_inv(key(0)); }
    template <typename T> __device__ double c1(const Prob<T>&) const { return dkey_pos_inv(key(1)); }
    template <typename T> __device__ void term_norms(const Prob<T>&, int, double& c2, double& fn) const {
        c2 = dkey_pos_inv(key(0));
        fn = dkey_pos_inv(key(1));
    }
    template <typename T> __device__ double trace_sum(const Prob<T>&, double*) const { return xr[4]; }
    template <typename T> __device__ void finish(const Prob<T>& P) const { P.ctrl->xepoch = epoch; }
};

// halo pushes: the same element of every peer's copy of an exchange-arena array.  The team is
// told, so that its next barrier orders these stores at system scope before it signals the peers.
template <class Team, typename V, typename T>
__device__ __forceinline__ void push_vec(const Team& team, const ShardDev& S, T* local, const V& v, unsigned mask) {
    if (mask) team.note_push();
    while (mask) {
        const int r = __ffs((int)mask) - 1;
        mask &= mask - 1u;
        v.store(reinterpret_cast<T*>(reinterpret_cast<char*>(local) + S.delta[r]));
    }
}
template <class Team>
__device__ __forceinline__ void push_f64(const Team& team, const ShardDev& S, double* local, double v, unsigned mask) {
    if (mask) team.note_push();
    while (mask) {
        const int r = __ffs((int)mask) - 1;
        mask &= mask - 1u;
        *reinterpret_cast<double*>(reinterpret_cast<char*>(local) + S.delta[r]) = v;
    }
}

// vector of VEC sketch columns held by one lane
template <typename T> struct Vec;
template <> struct Vec<double> {
    static constexpr int N = 2;
    double v[2];
    __device__ __forceinline__ void load(const double* p) {
        double2 t = *reinterpret_cast<const double2*>(p);
        v[0] = t.x; v[1] = t.y;
    }
    __device__ __forceinline__ void store(double* p) const { *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]); }
};
template <> struct Vec<float> {
    static constexpr int N = 4;
    float v[4];
    __device__ __forceinline__ void load(const float* p) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};


// ---------------------------------------------------------------------------
// TMA bulk copies (cp.async.bulk, SASS UBLKCP) tracked by an mbarrier
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// raw shared-memory loads by 32-bit shared address (keeps the address math 32-bit and the
// loads LDS, whatever the compiler can or cannot infer about the pointers)
template <typename T> struct SmemLd;
template <> struct SmemLd<double> {
    static __device__ __forceinline__ void vec(Vec<double>& v, unsigned a) {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.v[0]), "=d"(v.v[1]) : "r"(a));
    }
    static __device__ __forceinline__ double val(unsigned a) {
        double x;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a));
        return x;
    }
};
template <> struct SmemLd<float> {
    static __device__ __forceinline__ void vec(Vec<float>& v, unsigned a) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.v[0]), "=f"(v.v[1]), "=f"(v.v[2]), "=f"(v.v[3]) : "r"(a));
    }
    static __device__ __forceinline__ float val(unsigned a) {
        float x;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a));
        return x;
    }
};
__device__ __forceinline__ unsigned lds_u16(unsigned a) {
    unsigned short x;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(x) : "r"(a));
    return (unsigned)x;
}

// per-block staging state.  Dynamic shared memory layout (32-bit shared addresses):
//   [mbarrier 16 B][rows: ucap x Dp x T][vals: (nnzcap + 4) x T][lcol: (nnzcap + 8) u16]
template <typename T>
struct Stage {
    unsigned rows_a, vals_a, lcol_a, bar_a;
    unsigned parity;
    unsigned va, la;   // shared address of the staged value / local column of non-zero 0 (biased by the tile's first nnz)
};

__device__ __forceinline__ void mbar_expect_tx_a(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// try_wait suspends the thread for a hardware-defined time before it reports "not yet", so this
// loop is not a hot spin; debug builds (-DSIGSDP_DEBUG_SPIN) trap when a copy never lands
__device__ __forceinline__ void mbar_wait_a(unsigned bar, unsigned parity) {
    unsigned done;
#ifdef SIGSDP_DEBUG_SPIN
    unsigned tries = 0;
#endif
    do {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
#ifdef SIGSDP_DEBUG_SPIN
        if (!done && ++tries > (1u << 24)) __trap();
#endif
    } while (!done);
}

// Stage one tile with TMA bulk copies: the distinct rows of `src` (n x Dp, global) its
// non-zeros touch -- one copy per run of consecutive rows -- plus the tile's contiguous
// slices of `vals_src` (nnz values of the sketch dtype; nullptr to skip) and of the local
// column indices, each from a 16-byte aligned superset.  Completion is counted in bytes
// on the mbarrier; the compute that follows reads shared memory only.
template <typename T>
__device__ __forceinline__ void stage_tile(const Prob<T>& P, const T* src, const T* vals_src, int t, Stage<T>& st) {
    const TileDev& tl = P.tl;
    constexpr int VA = 16 / (int)sizeof(T);   // values per 16 bytes
    __syncthreads();  // every reader of the previous tile is done with the buffers
    const int4 ra = tl.trec[2 * t], rb = tl.trec[2 * t + 1];   // one load level instead of rowptr[trow[t]] chains
    const int p0 = ra.z, p1 = ra.w;
    const unsigned rowbytes = (unsigned)(P.Dp * sizeof(T));
    const int pv = p0 & ~(VA - 1), pl = p0 & ~7;
    const unsigned vbytes = vals_src ? (unsigned)(((p1 - pv + VA - 1) & ~(VA - 1)) * sizeof(T)) : 0u;
    const unsigned lbytes = (unsigned)(((p1 - pl + 7) & ~7) * 2);
    st.va = st.vals_a - (unsigned)(pv * (int)sizeof(T));
    st.la = st.lcol_a - (unsigned)(pl * 2);
    if (threadIdx.x == 0) {
        mbar_expect_tx_a(st.bar_a, (unsigned)rb.z * rowbytes + vbytes + lbytes);
        bulk_g2s_a(st.lcol_a, tl.lcol + pl, lbytes, st.bar_a);
        if (vals_src) bulk_g2s_a(st.vals_a, vals_src + pv, vbytes, st.bar_a);
    }
    // run i goes to warp i % NWARP, lane i / NWARP: a bulk copy is a per-warp instruction with
    // uniform operands (a warp issues its lanes' copies one after the other), so the runs
    // are dealt across the warps (and with them the SM's four schedulers) first
    for (int i = rb.x + (threadIdx.x >> 5) + NWARP * (threadIdx.x & 31); i < rb.y; i += NT) {
        const int4 r = tl.runs[i];
        bulk_g2s_a(st.rows_a + (unsigned)r.y * rowbytes, src + (size_t)r.x * P.Dp, (unsigned)r.z * rowbytes, st.bar_a);
    }
    mbar_wait_a(st.bar_a, st.parity);
    st.parity ^= 1u;
}

// ---------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller: counter-based normals for throughput mode
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}
// VEC standard normals for (iteration, row, vector index).  Box-Muller in fp32 for both
// sketch dtypes: the sketch only needs Gaussian directions, 24-bit normals are plenty.
__device__ __forceinline__ void philox_normals4(unsigned long long seed, long long iter, int row, int colv, float* out4) {
    uint4 x = philox4x32_10(make_uint4((unsigned)row, (unsigned)colv, (unsigned)iter, (unsigned)(iter >> 32)),
                            make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    float u1 = ((float)(x.x >> 8) + 0.5f) * 5.9604645e-8f, u2 = ((float)(x.y >> 8) + 0.5f) * 5.9604645e-8f;
    float u3 = ((float)(x.z >> 8) + 0.5f) * 5.9604645e-8f, u4 = ((float)(x.w >> 8) + 0.5f) * 5.9604645e-8f;
    float r1 = sqrtf(-2.0f * __logf(u1)), r2 = sqrtf(-2.0f * __logf(u3));
    float s, c;
    __sincosf(6.2831853f * u2, &s, &c);
    out4[0] = r1 * c;
    out4[1] = r1 * s;
    __sincosf(6.2831853f * u4, &s, &c);
    out4[2] = r2 * c;
    out4[3] = r2 * s;
}
__device__ __forceinline__ void philox_normals(unsigned long long seed, long long iter, int row, int colv, double* out2) {
    float t[4];
    philox_normals4(seed, iter, row, colv >> 1, t);   // two fp64 vectors share one Philox block
    out2[0] = (double)t[(colv & 1) * 2];
    out2[1] = (double)t[(colv & 1) * 2 + 1];
}
__device__ __forceinline__ void philox_normals(unsigned long long seed, long long iter, int row, int colv, float* out4) {
    philox_normals4(seed, iter, row, colv, out4);
}

// ---------------------------------------------------------------------------
// Taylor degree selection: scipy _fragment_3_1 (_expm_multiply.py:503-558), the branch
// taken when condition (3.13) holds: (m*, s) = argmin_m m ceil(||A||_1 / theta_m).
__constant__ double c_theta[35] = {2.29e-16, 2.58e-8, 1.39e-5, 3.40e-4, 2.40e-3, 9.07e-3, 2.38e-2, 5.00e-2, 8.96e-2,
                                   1.44e-1,  2.14e-1, 3.00e-1, 4.00e-1, 5.14e-1, 6.41e-1, 7.81e-1, 9.31e-1, 1.09,
                                   1.26,     1.44,    1.62,    1.82,    2.01,    2.22,    2.43,    2.64,    2.86,
                                   3.08,     3.31,    3.54,    4.7,     6.0,     7.2,     8.5,     9.9};
__constant__ int c_theta_m[35] = {1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18,
                                  19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 35, 40, 45, 50, 55};

__device__ __forceinline__ void taylor_select(double a1, int& m_star, long long& s) {
    if (!(a1 > 0.0)) {
        m_star = 0;
        s = 1;
        return;
    }
    double best = INFINITY;
    int bm = 1;
    double bs = 1.0;
#pragma unroll 1
    for (int i = 0; i < 35; ++i) {
        double si = ceil(a1 / c_theta[i]);
        double cost = (double)c_theta_m[i] * si;
        if (cost < best) {
            best = cost;
            bm = c_theta_m[i];
            bs = si;
        }
    }
    m_star = bm;
    s = bs < 9.0e15 ? (long long)bs : (long long)9.0e15;
}

// ===========================================================================
// Phase DUAL (mmw.py:124-139): e = [eD | eF | eH], e_accu += eta e, and the soft-max numerators.
//   eD_k = (X_kk - 1)/(1 - 1/K)
//   eF_e = (X_e + 1/(Z-1)) / (1/(K(Z-1)) + 1/2)            asso-UT edges
//   eH_k = ((T r)_k (Z-1)/Z - (h_k - S_sum_k/Z)) / norm_H_k,  r = row sums of X_offdiag
//          (quirk Q1: the reference's `*` is a sparse mat-mul)
// scipy.special.softmax (mmw.py:139) is exp(e - max e) / sum: ANY shift gives the same Y, so the
// numerators u = exp(e_accu - shift) are formed right here with the PREVIOUS iteration's maximum
// as the shift (e_accu moves by eta * e per iteration, so the exponent stays O(eta |e|) above
// zero) while the new maximum is reduced for the next iteration.  That removes a pass over the
// dual vector and one team barrier (a cross-GPU one when row-sharded) per iteration.  Outputs:
//   u, q_k = u_Hk / norm_H_k, partial sums S_D = sum u_D, S_F = sum u_F, S_H = sum u_H,
//   S_hq = sum_k hcoef_k q_k, and max e_accu.
template <typename T, int G, class Team>
__device__ void phase_dual(const Prob<T>& P, const Team& team, double shift, double* sh) {
    const PlanDev& g = P.g;
    const ShardDev& S = P.sh;
    const int K = g.n, Z = P.Z;
    cg::thread_block_tile<G> tile = cg::tiled_partition<G>(cg::this_thread_block());
    const int lane = tile.thread_rank();
    const int grp = threadIdx.x / G;
    constexpr int R = NT / G;
    const double invD = 1.0 / (1.0 - 1.0 / K);
    const double zr = (double)(Z - 1) / (double)Z;
    const double cF = 1.0 / ((double)K * (Z - 1)) + 0.5;
    double emax = -INFINITY;
    double sD = 0.0, sF = 0.0, sH = 0.0, sHq = 0.0;
    if (team.rank() == 0 && threadIdx.x < 3) {  // all of these are idle between gram and the loss phase
        P.ctrl->nrm_b[threadIdx.x] = 0ull;
        P.ctrl->nrm_f[threadIdx.x] = 0ull;
        if (threadIdx.x == 0) {
            P.ctrl->a1_key = 0ull;
            P.ctrl->c1_key = 0ull;
        }
    }
    const int ntiles = (S.row_hi - S.row_lo + R - 1) / R;
    for (int t = team.rank(); t < ntiles; t += team.size()) {
        const int k = S.row_lo + t * R + grp;
        if (k < S.row_hi) {
            // every load of the row is issued before the first use: the phase is bound by
            // memory latency, not bytes, so what counts is the number of loads in flight
            const int pa = g.rowptr[k], p1 = g.rowptr[k + 1];
            double hm = 0.0, ssum = 0.0, nh = 1.0, xd = 0.0, ea = 0.0, eb = 0.0, hc = 0.0;
            if (lane == 0) {
                hm = g.h_max[k];
                ssum = g.S_sum[k];
                nh = P.nH[k];
                hc = P.hcoef[k];
                xd = P.Xv[g.dpos[k]];
                ea = P.e_acc[k];
                eb = P.e_acc[K + g.E_a + k];
            }
            double acc = 0.0;
            for (int p = pa + lane; p < p1; p += 2 * G) {
                const bool hb = p + G < p1;
                const int pb = hb ? p + G : p;
                const double tf0 = g.tfwd[p], tf1 = g.tfwd[pb];
                const int c0 = g.col[p], c1 = g.col[pb];
                const double r0 = P.r[c0], r1 = P.r[c1];
                if (tf0 != 0.0) acc += tf0 * r0;
                if (hb && tf1 != 0.0) acc += tf1 * r1;
            }
            acc = group_sum<G>(tile, acc);
            if (lane == 0) {
                double eH = (acc * zr - (hm - ssum / Z)) / nh;
                double eD = (xd - 1.0) * invD;
                double a = ea + P.eta * eD;
                double b = eb + P.eta * eH;
                P.e_acc[k] = a;
                P.e_acc[K + g.E_a + k] = b;
                emax = fmax(emax, fmax(a, b));
                const double vD = exp(a - shift), vH = exp(b - shift);
                P.u[k] = vD;
                P.u[K + g.E_a + k] = vH;
                const double qq = vH / nh;
                P.q[k] = qq;
                if (S.pmask) push_f64(team, S, P.q + k, qq, S.pmask[k]);
                sD += vD;
                sH += vH;
                sHq += hc * qq;
            }
        }
    }
    // association edges with an entry in an own row (all of them when unsharded); a row-sharded
    // rank also advances the edges a neighbour owns from its own (bit-identical) copy of X_e
    const double zf = 1.0 / (Z - 1);
    for (int i = team.rank() * NT + threadIdx.x; i < S.n_inc; i += team.size() * NT) {
        const int e = S.inc_e ? S.inc_e[i] : i;
        const int pos = S.inc_e ? S.inc_pos[i] : g.apos[i];
        double eF = (P.Xv[pos] + zf) / cF;
        double a = P.e_acc[K + e] + P.eta * eF;
        P.e_acc[K + e] = a;
        const double v = exp(a - shift);
        P.u[K + e] = v;
        if (i < S.n_inc_owned) {
            emax = fmax(emax, a);
            sF += v;
        }
    }
    emax = block_max(emax, sh);
    sD = block_sum(sD, sh);
    sF = block_sum(sF, sh);
    sH = block_sum(sH, sh);
    sHq = block_sum(sHq, sh);
    if (threadIdx.x == 0) {
        if (emax > -INFINITY) atomicMax(&P.ctrl->emax_key, dkey_any(emax));
        double* o = P.psum + (size_t)team.rank() * PSTRIDE;
        o[0] = sD;
        o[1] = sF;
        o[2] = sH;
        o[3] = sHq;
    }
}

// ===========================================================================
// Phase LOSS (mmw.py:144-170 and csr_scal_rows_inplace, scipy_util.py:20-24, quirk Q2):
//   Y = u / S;  Y_avgd += Y_prev (mmw.py:78)
//   diag  l_k = (YD_k - sum YD / K)/(1 - 1/K) + (sum YF /(K(Z-1)))/cF - sum_k hcoef_k w_k
//   asso  l_e = (YF_e / 2)/cF
//   gain  l_e = (Z-1)/(2Z) (T[i,j] w_j + T[j,i] w_i),   w = YH / norm_H
//   L_accu -= eta l                                                      (mmw.py:167)
// plus what the sketch needs next: ||L_accu/2 - mu I||_1 with mu = tr(L_accu)/(2K)
// (scipy _expm_multiply.py:259-266), Omega_hat = rows of randn/sqrt(D) normalised
// (mmw.py:226-227) into B0 and F, ||Omega_hat||_inf, and ||Omega_hat_k||^2.
template <typename T, int G, class Team>
__device__ void phase_loss(const Prob<T>& P, const Team& team, int it_local, double* sh) {
    using V = Vec<T>;
    constexpr int VEC = V::N;
    const PlanDev& g = P.g;
    const ShardDev& S = P.sh;
    const int K = g.n, Z = P.Z, Dp = P.Dp, D = P.D;
    cg::thread_block_tile<G> tile = cg::tiled_partition<G>(cg::this_thread_block());
    const int lane = tile.thread_rank();
    const int grp = threadIdx.x / G;
    constexpr int R = NT / G;
    Ctrl* ctrl = P.ctrl;

    double es[4];
    team.exp_sums(P, sh, es);
    const double sD = es[0], sF = es[1], sH = es[2], sHq = es[3];
    const double Ssum = sD + sF + sH;
    const double cF = 1.0 / ((double)K * (Z - 1)) + 0.5;
    const double invD = 1.0 / (1.0 - 1.0 / K);
    const double sumYD = sD / Ssum;
    const double cLF = ((sF / Ssum) / ((double)K * (Z - 1))) / cF;
    const double cLH = sHq / Ssum;
    const double gcoef = (double)(Z - 1) / (2.0 * Z);
    const double eta = P.eta;
    const long long iter = ctrl->iter + it_local;
    // tr(L_accu): sum_k l_k = K (cLF - cLH) + (sum YD - sum YD)/(1-1/K) analytically
    const double trL = ctrl->trL[iter & 1] - eta * ((double)K * (cLF - cLH));
    const double mu = 0.5 * trL / K;
    if (team.rank() == 0 && threadIdx.x == 0) {
        ctrl->trL[(iter + 1) & 1] = trL;
        ctrl->mu = mu;
    }

    // ---- Y, Y_avgd (the entries this rank owns: own rows' D and H, owned asso edges' F)
    {
        const int nr = S.row_hi - S.row_lo;
        const int tot = 2 * nr + S.n_inc_owned;
        for (int i = team.rank() * NT + threadIdx.x; i < tot; i += team.size() * NT) {
            int c;
            if (i < nr) c = S.row_lo + i;
            else if (i < nr + S.n_inc_owned) c = K + (S.inc_e ? S.inc_e[i - nr] : i - nr);
            else c = K + g.E_a + S.row_lo + (i - nr - S.n_inc_owned);
            P.Ybar[c] += P.Y[c];
            P.Y[c] = P.u[c] / Ssum;
        }
    }

    // ---- L_accu, the shifted half A = L_accu/2 - mu I in the sketch dtype, and ||A||_1
    double a1 = 0.0;
    const double invS = 1.0 / Ssum;
    const double aF = invS * 0.5 / cF;
    const int ntiles = (S.row_hi - S.row_lo + R - 1) / R;
    for (int t = team.rank(); t < ntiles; t += team.size()) {
        const int k = S.row_lo + t * R + grp;
        if (k < S.row_hi) {
            const double wk = P.q[k] * invS;
            double rowabs = 0.0;
            const int p1 = g.rowptr[k + 1];
            const double ldiag = (P.u[k] * invS - sumYD / K) * invD + cLF - cLH;
            // two non-zeros per lane and trip, all their loads issued before the first use
            // (latency-bound phase: loads in flight are what counts)
            for (int p = g.rowptr[k] + lane; p < p1; p += 2 * G) {
                const bool hb = p + G < p1;
                const int pb = hb ? p + G : p;
                const int e0 = g.eid[p], e1 = g.eid[pb];
                const int c0 = g.col[p], c1 = g.col[pb];
                const double tf0 = g.tfwd[p], tf1 = g.tfwd[pb];
                const double tb0 = g.tbwd[p], tb1 = g.tbwd[pb];
                const double lv0 = P.Lval[p], lv1 = P.Lval[pb];
                const double q0 = P.q[c0], q1 = P.q[c1];
                const double ua0 = e0 >= g.E_g ? P.u[K + (e0 - g.E_g)] : 0.0;
                const double ua1 = e1 >= g.E_g ? P.u[K + (e1 - g.E_g)] : 0.0;
                {
                    double l, shift = 0.0;
                    if (e0 < 0) {
                        l = ldiag;
                        shift = mu;
                    } else if (e0 < g.E_g) {
                        l = gcoef * (tf0 * (q0 * invS) + tb0 * wk);
                    } else {
                        l = ua0 * aF;
                    }
                    const double v = lv0 - eta * l;
                    P.Lval[p] = v;
                    const double a = 0.5 * v - shift;
                    P.Aval[p] = (T)a;
                    rowabs += fabs(a);
                }
                if (hb) {
                    double l, shift = 0.0;
                    if (e1 < 0) {
                        l = ldiag;
                        shift = mu;
                    } else if (e1 < g.E_g) {
                        l = gcoef * (tf1 * (q1 * invS) + tb1 * wk);
                    } else {
                        l = ua1 * aF;
                    }
                    const double v = lv1 - eta * l;
                    P.Lval[pb] = v;
                    const double a = 0.5 * v - shift;
                    P.Aval[pb] = (T)a;
                    rowabs += fabs(a);
                }
            }
            rowabs = group_sum<G>(tile, rowabs);
            a1 = fmax(a1, rowabs);
        }
    }

    // ---- Omega_hat -> B0, F; ||.||_inf; ||row||^2
    double c1 = 0.0, trp = 0.0;
    const int Dtot = P.Dtot, col0 = P.col0;
    const double sqrtD = sqrt((double)Dtot);
    for (int t = team.rank(); t < ntiles; t += team.size()) {
        const int k = S.row_lo + t * R + grp;
        if (k < S.row_hi) {
            const int ko = g.perm ? g.perm[k] : k;
            const unsigned pm = S.pmask ? S.pmask[k] : 0u;
            double ss = 0.0;
            // the row norm runs over all Dtot columns of the sketch: a column shard generates
            // (or reads) the columns it does not own only for that
            if (Dtot != D)
                for (int c0 = lane * VEC; c0 < Dtot; c0 += G * VEC) {
                    if (c0 >= col0 && c0 < col0 + D) continue;
                    T raw[VEC];
                    if (P.omega) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v)
                            raw[v] = (c0 + v < Dtot) ? (T)(P.omega[((size_t)it_local * K + ko) * Dtot + c0 + v] / sqrtD) : (T)0;
                    } else {
                        philox_normals(P.seed, iter, ko, c0 / VEC, raw);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) raw[v] = (c0 + v < Dtot) ? raw[v] : (T)0;   // 1/sqrt(D) cancels below
                    }
#pragma unroll
                    for (int v = 0; v < VEC; ++v) ss += (double)raw[v] * (double)raw[v];
                }
            const bool one = Dp == G * VEC;
            V w0;
#pragma unroll
            for (int v = 0; v < VEC; ++v) w0.v[v] = (T)0;
            for (int c0 = lane * VEC; c0 < Dp; c0 += G * VEC) {
                T raw[VEC];
                if (P.omega) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
                        raw[v] = (c0 + v < D) ? (T)(P.omega[((size_t)it_local * K + ko) * Dtot + col0 + c0 + v] / sqrtD) : (T)0;
                } else {
                    philox_normals(P.seed, iter, ko, (col0 + c0) / VEC, raw);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) raw[v] = (c0 + v < D) ? raw[v] : (T)0;
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) ss += (double)raw[v] * (double)raw[v];
#pragma unroll
                for (int v = 0; v < VEC; ++v) w0.v[v] = raw[v];
                if (!one) w0.store(P.B0 + (size_t)k * Dp + c0);  // unnormalised for now
            }
            ss = group_sum<G>(tile, ss);
            // Omega from the caller: the reference's arithmetic (randn / sqrt(D), row / its norm).
            // Device generator: the 1/sqrt(D) cancels and one reciprocal square root scales the row.
            const bool host_omega = P.omega != nullptr;
            const double nrm = host_omega ? sqrt(ss) : 1.0;
            const double inrm = host_omega ? 1.0 : rsqrt(ss);
            double rs = 0.0, dd = 0.0;
            for (int c0 = lane * VEC; c0 < Dp; c0 += G * VEC) {
                V w = w0;   // one chunk per lane: the row never leaves the registers
                if (!one) w.load(P.B0 + (size_t)k * Dp + c0);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    w.v[v] = host_omega ? (T)((double)w.v[v] / nrm) : (T)((double)w.v[v] * inrm);
                    rs += fabs((double)w.v[v]);
                    dd += (double)w.v[v] * (double)w.v[v];
                }
                w.store(P.B0 + (size_t)k * Dp + c0);
                w.store(P.F + (size_t)k * Dp + c0);
                if (pm) {
                    push_vec(team, S, P.B0 + (size_t)k * Dp + c0, w, pm);
                    push_vec(team, S, P.F + (size_t)k * Dp + c0, w, pm);
                }
            }
            rs = group_sum<G>(tile, rs);
            dd = group_sum<G>(tile, dd);
            c1 = fmax(c1, rs);
            if (lane == 0) {
                P.dsq[k] = dd;
                trp += dd;
            }
        }
    }
    a1 = block_max(a1, sh);
    c1 = block_max(c1, sh);
    trp = block_sum(trp, sh);
    if (threadIdx.x == 0) {
        atomicMax(&ctrl->a1_key, dkey_pos(a1));
        atomicMax(&ctrl->c1_key, dkey_pos(c1));
        P.ptr[team.rank()] = trp;
    }
}

// ===========================================================================
// One Taylor term (scipy _expm_multiply_simple_core, _expm_multiply.py:291-303), fused:
//   B_new = coeff (A - mu I) B,  F += B_new,  ||B_new||_inf,  ||F||_inf,  ||F_k||^2
// with A = L_accu / 2 read straight from the fp64 accumulator.  One group of G lanes
// per row, each lane holds VEC consecutive sketch columns (16-byte loads); column
// indices and values of G non-zeros are loaded coalesced and broadcast by shuffle.
template <typename T, int G, class Team>
__device__ void phase_term(const Prob<T>& P, const Team& team, const T* Bin, T* Bout, double coeff, double mu,
                           int slot, double* sh) {
    using V = Vec<T>;
    constexpr int VEC = V::N;
    const PlanDev& g = P.g;
    const ShardDev& S = P.sh;
    const int Dp = P.Dp;
    cg::thread_block_tile<G> tile = cg::tiled_partition<G>(cg::this_thread_block());
    const int lane = tile.thread_rank();
    const int grp = threadIdx.x / G;
    constexpr int R = NT / G;
    Ctrl* ctrl = P.ctrl;
    if (team.rank() == 0 && threadIdx.x == 0) {  // slot+1 was last read two barriers ago
        ctrl->nrm_b[(slot + 1) % 3] = 0ull;
        ctrl->nrm_f[(slot + 1) % 3] = 0ull;
    }
    const T cf = (T)coeff;
    double bmax = 0.0, fmaxv = 0.0, trp = 0.0;
    const int ntiles = (S.row_hi - S.row_lo + R - 1) / R;
    for (int t = team.rank(); t < ntiles; t += team.size()) {
        const int k = S.row_lo + t * R + grp;
        if (k < S.row_hi) {
            const int p0 = g.rowptr[k], p1 = g.rowptr[k + 1];
            const unsigned pm = S.pmask ? S.pmask[k] : 0u;
            double rsb = 0.0, rsf = 0.0, dd = 0.0;
            for (int cb = 0; cb < Dp; cb += G * VEC) {
                // every lane of the group runs every chunk so the shuffles below stay
                // convergent; lanes past the row end only broadcast
                const int c0 = cb + lane * VEC;
                const bool act = c0 < Dp;
                T acc[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v] = (T)0;
                for (int pb = p0; pb < p1; pb += G) {
                    const int p = pb + lane;
                    int cc = 0;
                    T vv = (T)0;
                    if (p < p1) {
                        cc = g.col[p];
                        vv = P.Aval[p];
                    }
                    const int cnt = min(G, p1 - pb);
#pragma unroll 4
                    for (int j = 0; j < cnt; ++j) {
                        const int cj = tile.shfl(cc, j);
                        const T vj = tile.shfl(vv, j);
                        if (act) {
                            V b;
                            b.load(Bin + (size_t)cj * Dp + c0);
#pragma unroll
                            for (int v = 0; v < VEC; ++v) acc[v] = fma(vj, b.v[v], acc[v]);
                        }
                    }
                }
                if (act) {
                    V bn, f;
                    f.load(P.F + (size_t)k * Dp + c0);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        bn.v[v] = cf * acc[v];
                        f.v[v] += bn.v[v];
                        rsb += fabs((double)bn.v[v]);
                        rsf += fabs((double)f.v[v]);
                        dd += (double)f.v[v] * (double)f.v[v];
                    }
                    bn.store(Bout + (size_t)k * Dp + c0);
                    f.store(P.F + (size_t)k * Dp + c0);
                    if (pm) {
                        push_vec(team, S, Bout + (size_t)k * Dp + c0, bn, pm);
                        push_vec(team, S, P.F + (size_t)k * Dp + c0, f, pm);
                    }
                }
            }
            rsb = group_sum<G>(tile, rsb);
            rsf = group_sum<G>(tile, rsf);
            dd = group_sum<G>(tile, dd);
            bmax = fmax(bmax, rsb);
            fmaxv = fmax(fmaxv, rsf);
            if (lane == 0) {
                P.dsq[k] = dd;
                trp += dd;
            }
        }
    }
    bmax = block_max(bmax, sh);
    fmaxv = block_max(fmaxv, sh);
    trp = block_sum(trp, sh);
    if (threadIdx.x == 0) {
        atomicMax(&ctrl->nrm_b[slot], dkey_pos(bmax));
        atomicMax(&ctrl->nrm_f[slot], dkey_pos(fmaxv));
        P.ptr[team.rank()] = trp;
    }
}

// B <- F at an s-step boundary (scipy _expm_multiply.py:301-303 without the eta scaling,
// which is deferred: exp(A) B = e^mu prod_s T_m((A - mu I)/s) B)
template <typename T, int G, class Team>
__device__ void phase_copy(const Prob<T>& P, const Team& team, T* Bdst) {
    const size_t tot = (size_t)P.g.n * P.Dp;
    for (size_t i = (size_t)team.rank() * NT + threadIdx.x; i < tot; i += (size_t)team.size() * NT) Bdst[i] = P.F[i];
}

// ===========================================================================
// Phase GRAM (mmw.py:182-194, 77): X = sketch Gram on the pattern only, normalised by
// tr = sum_k ||y_k||^2 / K; running sum X_avgd += X_prev; r = row sums of X_offdiag for
// the next dual step.  Row-parallel over the symmetric pattern: each directed entry is
// a dot product of two sketch rows; the entry with row < col owns the edge's storage.
template <typename T, int G, class Team>
__device__ void phase_gram(const Prob<T>& P, const Team& team, double* sh) {
    using V = Vec<T>;
    constexpr int VEC = V::N;
    const PlanDev& g = P.g;
    const ShardDev& S = P.sh;
    const int K = g.n, Dp = P.Dp;
    cg::thread_block_tile<G> tile = cg::tiled_partition<G>(cg::this_thread_block());
    const int lane = tile.thread_rank();
    const int grp = threadIdx.x / G;
    constexpr int R = NT / G;
    const double tr = P.split ? 1.0 : team.trace_sum(P, sh) / K;
    const bool one_chunk = Dp <= G * VEC;
    const int ntiles = (S.row_hi - S.row_lo + R - 1) / R;
    for (int t = team.rank(); t < ntiles; t += team.size()) {
        const int k = S.row_lo + t * R + grp;
        if (k < S.row_hi) {
            const int p0 = g.rowptr[k], p1 = g.rowptr[k + 1];
            V fk;
#pragma unroll
            for (int v = 0; v < VEC; ++v) fk.v[v] = (T)0;
            if (one_chunk && lane * VEC < Dp) fk.load(P.F + (size_t)k * Dp + lane * VEC);
            double rsum = 0.0;
            for (int pb = p0; pb < p1; pb += G) {
                const int p = pb + lane;
                int cc = 0, ee = -1;
                if (p < p1) {
                    cc = g.col[p];
                    ee = g.eid[p];
                }
                const int cnt = min(G, p1 - pb);
                double mine = 0.0;
                for (int j = 0; j < cnt; ++j) {
                    const int cj = tile.shfl(cc, j);
                    T part = (T)0;
                    if (cj != k) {
                        if (one_chunk) {
                            if (lane * VEC < Dp) {
                                V b;
                                b.load(P.F + (size_t)cj * Dp + lane * VEC);
#pragma unroll
                                for (int v = 0; v < VEC; ++v) part = fma(fk.v[v], b.v[v], part);
                            }
                        } else {
                            for (int c0 = lane * VEC; c0 < Dp; c0 += G * VEC) {
                                V a, b;
                                a.load(P.F + (size_t)k * Dp + c0);
                                b.load(P.F + (size_t)cj * Dp + c0);
#pragma unroll
                                for (int v = 0; v < VEC; ++v) part = fma(a.v[v], b.v[v], part);
                            }
                        }
                    }
                    double dot = group_sum<G>(tile, (double)part);
                    if (j == lane) mine = dot;
                }
                if (p < p1) {
                    if (P.split) {
                        P.graw[p] = ee >= 0 ? mine : 0.0;
                    } else {
                        const double x = ee >= 0 ? mine / tr : P.dsq[k] / tr;
                        if (ee >= 0) rsum += x;
                        P.Xbarv[p] += P.Xv[p];
                        P.Xv[p] = x;
                    }
                }
            }
            rsum = group_sum<G>(tile, rsum);
            if (lane == 0 && !P.split) {
                P.r[k] = rsum;
                if (S.pmask) push_f64(team, S, P.r + k, rsum, S.pmask[k]);
            }
        }
    }
}


// ===========================================================================
// Staged Taylor term: same arithmetic as phase_term, but everything a tile needs -- the
// distinct rows of B its non-zeros touch, its slice of A = L_accu/2 - mu I and its local
// column indices -- is brought into shared memory by TMA bulk copies first, so the multiply
// loop never waits on global memory: per non-zero one 2-byte and one value broadcast load,
// one 16-byte row load and VEC FMAs.  L2 -> SM traffic for B drops by the tile's reuse
// factor, and the two resident blocks per SM overlap one tile's copies with the other's math.
template <typename T, int G, class Team>
__device__ void phase_term_staged(const Prob<T>& P, const Team& team, const T* Bin, T* Bout, double coeff, int slot,
                                  double* sh, Stage<T>& st) {
    using V = Vec<T>;
    using LD = SmemLd<T>;
    constexpr int VEC = V::N;
    constexpr int W = (int)sizeof(T);
    const PlanDev& g = P.g;
    const TileDev& tl = P.tl;
    const int Dp = P.Dp;
    const unsigned rowb = (unsigned)(Dp * W);
    const int lane = threadIdx.x & (G - 1);
    const int grp = threadIdx.x / G;
    constexpr int R = NT / G;
    Ctrl* ctrl = P.ctrl;
    if (team.rank() == 0 && threadIdx.x == 0) {
        ctrl->nrm_b[(slot + 1) % 3] = 0ull;
        ctrl->nrm_f[(slot + 1) % 3] = 0ull;
    }
    const T cf = (T)coeff;
    double bmax = 0.0, fmaxv = 0.0, trp = 0.0;
    const ShardDev& S = P.sh;
    fence_proxy_async();   // order this phase's bulk copies after the barrier that published their source
    for (int t = S.tile_lo + team.rank(); t < S.tile_hi; t += team.size()) {
        const int4 trc = tl.trec[2 * t];
        const int r0 = trc.x, r1 = trc.y;
        stage_tile(P, Bin, (const T*)P.Aval, t, st);
        for (int kb = r0; kb < r1; kb += R) {   // block-uniform trip count
            const int k = kb + grp;
            const bool valid = k < r1;
            const int p0 = valid ? g.rowptr[k] : 0, len = valid ? g.rowptr[k + 1] - p0 : 0;
            const unsigned pm = (valid && S.pmask) ? S.pmask[k] : 0u;
            double rsb = 0.0, rsf = 0.0, dd = 0.0;
            for (int cb = 0; cb < Dp; cb += G * VEC) {
                const int c0 = cb + lane * VEC;
                const bool act = valid && c0 < Dp;
                V f;
#pragma unroll
                for (int v = 0; v < VEC; ++v) f.v[v] = (T)0;
                if (act) f.load(P.F + (size_t)k * Dp + c0);   // consumed in the epilogue
                T acc[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v] = (T)0;
                if (act) {
                    const unsigned la = st.la + 2u * (unsigned)p0, va = st.va + (unsigned)(W * p0);
                    const unsigned rb = st.rows_a + (unsigned)(c0 * W);
                    int j = 0;
                    for (; j + 4 <= len; j += 4) {
                        unsigned lc[4];
                        T a[4];
                        V b[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            lc[i] = lds_u16(la + 2u * (unsigned)(j + i));
                            a[i] = LD::val(va + (unsigned)(W * (j + i)));
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) LD::vec(b[i], rb + lc[i] * rowb);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int v = 0; v < VEC; ++v) acc[v] = fma(a[i], b[i].v[v], acc[v]);
                    }
                    for (; j < len; ++j) {
                        const unsigned lc = lds_u16(la + 2u * (unsigned)j);
                        const T a = LD::val(va + (unsigned)(W * j));
                        V b;
                        LD::vec(b, rb + lc * rowb);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) acc[v] = fma(a, b.v[v], acc[v]);
                    }
                    V bn;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        bn.v[v] = cf * acc[v];
                        f.v[v] += bn.v[v];
                        rsb += fabs((double)bn.v[v]);
                        rsf += fabs((double)f.v[v]);
                        dd += (double)f.v[v] * (double)f.v[v];
                    }
                    bn.store(Bout + (size_t)k * Dp + c0);
                    f.store(P.F + (size_t)k * Dp + c0);
                    if (pm) {
                        push_vec(team, S, Bout + (size_t)k * Dp + c0, bn, pm);
                        push_vec(team, S, P.F + (size_t)k * Dp + c0, f, pm);
                    }
                }
            }
            // all 32 lanes are converged here: xor-shuffles below G stay inside the group
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                rsb += __shfl_xor_sync(0xffffffffu, rsb, o);
                rsf += __shfl_xor_sync(0xffffffffu, rsf, o);
                dd += __shfl_xor_sync(0xffffffffu, dd, o);
            }
            bmax = fmax(bmax, rsb);
            fmaxv = fmax(fmaxv, rsf);
            if (valid && lane == 0) {
                P.dsq[k] = dd;
                trp += dd;
            }
        }
    }
    bmax = block_max(bmax, sh);
    fmaxv = block_max(fmaxv, sh);
    trp = block_sum(trp, sh);
    if (threadIdx.x == 0) {
        atomicMax(&ctrl->nrm_b[slot], dkey_pos(bmax));
        atomicMax(&ctrl->nrm_f[slot], dkey_pos(fmaxv));
        P.ptr[team.rank()] = trp;
    }
}


// Variant of phase_term_staged for rows that fill the lane group exactly (Dp == G * VEC,
// G >= 8): G/2 lanes per row, each lane owns TWO 16-byte chunks of the row (columns
// [lane*VEC, +VEC) and [(lane + G/2)*VEC, +VEC), so each of its two loads is conflict-free
// across the group).  The per-non-zero index / value broadcasts are then shared by twice as
// many columns: 2.5 instead of 3 shared-memory wavefronts and ~30 % fewer instructions per
// non-zero, and a 64-row tile is one pass of a 512-thread block.
template <typename T, int G, class Team>
__device__ void phase_term_staged2(const Prob<T>& P, const Team& team, const T* Bin, T* Bout, double coeff, int slot,
                                   double* sh, Stage<T>& st) {
    using V = Vec<T>;
    using LD = SmemLd<T>;
    constexpr int VEC = V::N;
    constexpr int W = (int)sizeof(T);
    constexpr int GH = G / 2;
    constexpr int R = NT / GH;
    const PlanDev& g = P.g;
    const TileDev& tl = P.tl;
    const ShardDev& S = P.sh;
    const int Dp = P.Dp;
    const unsigned rowb = (unsigned)(Dp * W);
    const int lane = threadIdx.x & (GH - 1);
    const int grp = threadIdx.x / GH;
    Ctrl* ctrl = P.ctrl;
    if (team.rank() == 0 && threadIdx.x == 0) {
        ctrl->nrm_b[(slot + 1) % 3] = 0ull;
        ctrl->nrm_f[(slot + 1) % 3] = 0ull;
    }
    const T cf = (T)coeff;
    double bmax = 0.0, fmaxv = 0.0, trp = 0.0;
    const int ca = lane * VEC, cb2 = (lane + GH) * VEC;
    const bool use_slots = tl.slots != nullptr && tl.slot_r == R;
    fence_proxy_async();
    for (int t = S.tile_lo + team.rank(); t < S.tile_hi; t += team.size()) {
        const int4 trc = tl.trec[2 * t];
        const int r0 = trc.x, r1 = trc.y;
        // this group's slot record is requested before the stage: its latency (and that of the
        // F row it names) hides behind the bulk copies instead of following them
        int4 sr = make_int4(-1, 0, 0, 0);
        if (use_slots) sr = tl.slots[(size_t)t * R + grp];
        stage_tile(P, Bin, (const T*)P.Aval, t, st);
        for (int kb = r0; kb < r1; kb += R) {   // block-uniform trip count (one trip with a slot table)
            int k, p0, len, role = 0;
            bool valid;
            if (use_slots) {
                k = sr.x;
                valid = k >= 0;
                p0 = sr.y;
                len = sr.z & 0xffff;
                role = sr.z >> 16;
            } else {
                k = kb + grp;
                valid = k < r1;
                p0 = valid ? g.rowptr[k] : 0;
                len = valid ? g.rowptr[k + 1] - p0 : 0;
            }
            const bool owner = valid && role != 2;
            double rsb = 0.0, rsf = 0.0, dd = 0.0;
            V fa, fb;
            T acca[VEC], accb[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) acca[v] = accb[v] = fa.v[v] = fb.v[v] = (T)0;
            unsigned pm = 0u;
            if (owner) {
                fa.load(P.F + (size_t)k * Dp + ca);   // consumed in the epilogue
                fb.load(P.F + (size_t)k * Dp + cb2);
                if (S.pmask) pm = S.pmask[k];
            }
            if (valid) {
                const unsigned la = st.la + 2u * (unsigned)p0, va = st.va + (unsigned)(W * p0);
                const unsigned rba = st.rows_a + (unsigned)(ca * W), rbb = st.rows_a + (unsigned)(cb2 * W);
                int j = 0;
                for (; j + 4 <= len; j += 4) {
                    unsigned lc[4];
                    T a[4];
                    V ba[4], bb[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        lc[i] = lds_u16(la + 2u * (unsigned)(j + i)) * rowb;
                        a[i] = LD::val(va + (unsigned)(W * (j + i)));
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        LD::vec(ba[i], rba + lc[i]);
                        LD::vec(bb[i], rbb + lc[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            acca[v] = fma(a[i], ba[i].v[v], acca[v]);
                            accb[v] = fma(a[i], bb[i].v[v], accb[v]);
                        }
                }
                for (; j < len; ++j) {
                    const unsigned lc = lds_u16(la + 2u * (unsigned)j) * rowb;
                    const T a = LD::val(va + (unsigned)(W * j));
                    V ba, bb;
                    LD::vec(ba, rba + lc);
                    LD::vec(bb, rbb + lc);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        acca[v] = fma(a, ba.v[v], acca[v]);
                        accb[v] = fma(a, bb.v[v], accb[v]);
                    }
                }
            }
            // second halves hand their partial sums to the group on their left (same warp: pairs
            // start at even slots); a warp without split rows skips the shuffles
            if (__any_sync(0xffffffffu, role == 1)) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const T ra = __shfl_down_sync(0xffffffffu, acca[v], GH);
                    const T rb = __shfl_down_sync(0xffffffffu, accb[v], GH);
                    if (role == 1) {
                        acca[v] += ra;
                        accb[v] += rb;
                    }
                }
            }
            if (owner) {
                V bna, bnb;
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    bna.v[v] = cf * acca[v];
                    bnb.v[v] = cf * accb[v];
                    fa.v[v] += bna.v[v];
                    fb.v[v] += bnb.v[v];
                    rsb += fabs((double)bna.v[v]) + fabs((double)bnb.v[v]);
                    rsf += fabs((double)fa.v[v]) + fabs((double)fb.v[v]);
                    dd += (double)fa.v[v] * (double)fa.v[v] + (double)fb.v[v] * (double)fb.v[v];
                }
                bna.store(Bout + (size_t)k * Dp + ca);
                bnb.store(Bout + (size_t)k * Dp + cb2);
                fa.store(P.F + (size_t)k * Dp + ca);
                fb.store(P.F + (size_t)k * Dp + cb2);
                if (pm) {   // rows a neighbouring rank reads: the same stores into its copies (NVLink)
                    push_vec(team, S, Bout + (size_t)k * Dp + ca, bna, pm);
                    push_vec(team, S, Bout + (size_t)k * Dp + cb2, bnb, pm);
                    push_vec(team, S, P.F + (size_t)k * Dp + ca, fa, pm);
                    push_vec(team, S, P.F + (size_t)k * Dp + cb2, fb, pm);
                }
            }
            // all 32 lanes are converged here: xor-shuffles below GH stay inside the group
#pragma unroll
            for (int o = GH / 2; o > 0; o >>= 1) {
                rsb += __shfl_xor_sync(0xffffffffu, rsb, o);
                rsf += __shfl_xor_sync(0xffffffffu, rsf, o);
                dd += __shfl_xor_sync(0xffffffffu, dd, o);
            }
            bmax = fmax(bmax, rsb);
            fmaxv = fmax(fmaxv, rsf);
            if (owner && lane == 0) {
                P.dsq[k] = dd;
                trp += dd;
            }
        }
    }
    bmax = block_max(bmax, sh);
    fmaxv = block_max(fmaxv, sh);
    trp = block_sum(trp, sh);
    if (threadIdx.x == 0) {
        atomicMax(&ctrl->nrm_b[slot], dkey_pos(bmax));
        atomicMax(&ctrl->nrm_f[slot], dkey_pos(fmaxv));
        P.ptr[team.rank()] = trp;
    }
}

// Staged Gram: the tile's rows of F (its own rows included: the diagonal is in the
// pattern) and its local column indices sit in shared memory; one warp per row, one lane
// per non-zero, each lane accumulates its own dot product over the sketch columns.  No
// shuffles; each lane starts at a different 16-byte chunk and wraps around, so the eight
// lanes of a quarter-warp always hit eight different bank groups whatever rows they read.
template <typename T, int G, class Team>
__device__ void phase_gram_staged(const Prob<T>& P, const Team& team, double* sh, Stage<T>& st) {
    using V = Vec<T>;
    using LD = SmemLd<T>;
    constexpr int VEC = V::N;
    constexpr int W = (int)sizeof(T);
    const PlanDev& g = P.g;
    const TileDev& tl = P.tl;
    const ShardDev& S = P.sh;
    const int K = g.n, Dp = P.Dp;
    const unsigned rowb = (unsigned)(Dp * W);
    const int nc = Dp / VEC;
    const bool pow2 = (rowb & (rowb - 1u)) == 0u && nc >= 2;
    const unsigned rmask = rowb - 1u;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const bool raw = P.split != 0;
    const double tr = raw ? 1.0 : team.trace_sum(P, sh) / K;
    const double inv_tr = 1.0 / tr;
    fence_proxy_async();
    for (int t = S.tile_lo + team.rank(); t < S.tile_hi; t += team.size()) {
        const int4 trc = tl.trec[2 * t];
        const int r0 = trc.x, r1 = trc.y;
        // Row pointers / diagonal positions of this warp's rows (lane i: its i-th row of the
        // tile) are requested before the stage and handed out by shuffle: one global latency per
        // tile, hidden behind the stage, instead of one per row.
        int mp0 = 0, mp1 = 0, mpd = 0;
        {
            const int kk = r0 + wrp + NWARP * lane;
            if (kk < r1) {
                mp0 = g.rowptr[kk];
                mp1 = g.rowptr[kk + 1];
                mpd = g.dpos[kk];
            }
        }
        stage_tile(P, (const T*)P.F, (const T*)nullptr, t, st);
        int pass = 0;
        for (int kb = r0; kb < r1; kb += NWARP, ++pass) {   // block-uniform trip count
            const int k = kb + wrp;
            double rsum = 0.0;
            int p0, p1, pd;
            if (pass < 32) {
                p0 = __shfl_sync(0xffffffffu, mp0, pass);
                p1 = __shfl_sync(0xffffffffu, mp1, pass);
                pd = __shfl_sync(0xffffffffu, mpd, pass);
            } else {   // tiles of more than 32 * NWARP rows do not exist today
                p0 = k < r1 ? g.rowptr[k] : 0;
                p1 = k < r1 ? g.rowptr[k + 1] : 0;
                pd = k < r1 ? g.dpos[k] : 0;
            }
            if (k < r1) {
                const unsigned abase = st.rows_a + lds_u16(st.la + 2u * (unsigned)pd) * rowb;
                for (int p = p0 + lane; p < p1; p += 32) {
                    double xold = 0.0, xbar = 0.0;
                    if (!raw) {   // in flight during the dot product
                        xold = P.Xv[p];
                        xbar = P.Xbarv[p];
                    }
                    double x;
                    if (p == pd) {
                        x = raw ? 0.0 : P.dsq[k] * inv_tr;
                    } else {
                        const unsigned bbase = st.rows_a + lds_u16(st.la + 2u * (unsigned)p) * rowb;
                        T d0 = (T)0, d1 = (T)0;
                        if (pow2) {   // row bytes are a power of two: the rotation is a mask
                            unsigned off = (unsigned)(lane * 16) & rmask;
#pragma unroll 2
                            for (int s2 = 0; s2 < nc; s2 += 2) {
                                V a0, b0, a1, b1;
                                const unsigned off1 = (off + 16u) & rmask;
                                LD::vec(a0, abase + off);
                                LD::vec(b0, bbase + off);
                                LD::vec(a1, abase + off1);
                                LD::vec(b1, bbase + off1);
                                off = (off1 + 16u) & rmask;
#pragma unroll
                                for (int v = 0; v < VEC; ++v) {
                                    d0 = fma(a0.v[v], b0.v[v], d0);
                                    d1 = fma(a1.v[v], b1.v[v], d1);
                                }
                            }
                        } else {
                        int ch = lane % nc;
                        for (int s2 = 0; s2 < nc; s2 += 2) {
                            V a, b;
                            unsigned off = (unsigned)(ch * 16);
                            LD::vec(a, abase + off);
                            LD::vec(b, bbase + off);
#pragma unroll
                            for (int v = 0; v < VEC; ++v) d0 = fma(a.v[v], b.v[v], d0);
                            ch = ch + 1 == nc ? 0 : ch + 1;
                            if (s2 + 1 < nc) {
                                off = (unsigned)(ch * 16);
                                LD::vec(a, abase + off);
                                LD::vec(b, bbase + off);
#pragma unroll
                                for (int v = 0; v < VEC; ++v) d1 = fma(a.v[v], b.v[v], d1);
                                ch = ch + 1 == nc ? 0 : ch + 1;
                            }
                        }
                        }
                        x = ((double)d0 + (double)d1) * inv_tr;
                        rsum += x;
                    }
                    if (raw) {
                        P.graw[p] = x;
                    } else {
                        P.Xbarv[p] = xbar + xold;
                        P.Xv[p] = x;
                    }
                }
            }
            rsum = warp_sum(rsum);
            if (k < r1 && lane == 0 && !raw) {
                P.r[k] = rsum;
                if (S.pmask) push_f64(team, S, P.r + k, rsum, S.pmask[k]);
            }
        }
    }
}


// Variant of phase_gram_staged for rows that fill the lane group exactly (Dp == G * VEC,
// G >= 8), same thread mapping as phase_term_staged2: G/2 lanes per row, each lane keeps ITS
// two 16-byte chunks of F_k in registers for the whole row and reads the same two chunks of
// every neighbour row F_c, i.e. 2 shared-memory wavefronts per non-zero instead of the 4-5 of
// the lane-per-non-zero kernel (which re-reads F_k per non-zero).  The G/2 partial dot
// products of a non-zero are combined eight non-zeros at a time by a transposing butterfly
// (xor 4, 2, 1: 7 shuffles for 8 sums, lane l ends up owning non-zero l of the chunk) and
// that lane finishes the entry (X, X_avgd, row sum).  Fixed summation tree: bit-reproducible.
template <typename T, int G, class Team>
__device__ void phase_gram_staged2(const Prob<T>& P, const Team& team, double* sh, Stage<T>& st) {
    using V = Vec<T>;
    using LD = SmemLd<T>;
    constexpr int VEC = V::N;
    constexpr int W = (int)sizeof(T);
    constexpr int GH = G / 2;
    constexpr int R = NT / GH;
    const PlanDev& g = P.g;
    const TileDev& tl = P.tl;
    const ShardDev& S = P.sh;
    const int K = g.n, Dp = P.Dp;
    const unsigned rowb = (unsigned)(Dp * W);
    const int lane = threadIdx.x & (GH - 1);
    const int grp = threadIdx.x / GH;
    const bool raw = P.split != 0;
    const double tr = raw ? 1.0 : team.trace_sum(P, sh) / K;
    const double inv_tr = 1.0 / tr;
    const unsigned ca = (unsigned)(lane * VEC * W), cb2 = (unsigned)((lane + GH) * VEC * W);
    const bool use_slots = tl.slots != nullptr && tl.slot_r == R;
    fence_proxy_async();
    for (int t = S.tile_lo + team.rank(); t < S.tile_hi; t += team.size()) {
        const int4 trc = tl.trec[2 * t];
        const int r0 = trc.x, r1 = trc.y;
        int4 sr = make_int4(-1, 0, 0, 0);   // requested before the stage (see phase_term_staged2)
        if (use_slots) sr = tl.slots[(size_t)t * R + grp];
        stage_tile(P, (const T*)P.F, (const T*)nullptr, t, st);
        for (int kb = r0; kb < r1; kb += R) {   // block-uniform trip count (one trip with a slot table)
            // the term kernel's slot table balances this kernel too: the two halves of a split
            // row finish their own entries independently, only the row sum is combined
            int k, p0 = 0, len = 0, pd = 0, role = 0;
            bool valid;
            if (use_slots) {
                k = sr.x;
                valid = k >= 0;
                p0 = sr.y;
                len = sr.z & 0xffff;
                role = sr.z >> 16;
                pd = sr.w;   // position of the row's diagonal entry
            } else {
                k = kb + grp;
                valid = k < r1;
                if (valid) {
                    p0 = g.rowptr[k];
                    len = g.rowptr[k + 1] - p0;
                    pd = g.dpos[k];
                }
            }
            // F_k: this lane's two chunks (an idle group reads slot 0 of the tile and discards)
            V fka, fkb;
            {
                const unsigned abase = st.rows_a + (valid ? lds_u16(st.la + 2u * (unsigned)pd) * rowb : 0u);
                LD::vec(fka, abase + ca);
                LD::vec(fkb, abase + cb2);
            }
            const double xdiag = (valid && !raw) ? P.dsq[k] * inv_tr : 0.0;
            const unsigned la = st.la + 2u * (unsigned)p0;
            // the shuffles below are warp-wide instructions: every group of the warp runs the
            // trip count of the warp's longest row (shorter rows repeat their last entry)
            const int nch = __reduce_max_sync(0xffffffffu, (len + GH - 1) / GH);
            double rsum = 0.0;
            for (int c = 0; c < nch; ++c) {
                const int jmine = c * GH + lane;           // the entry this lane will finish
                const bool mine = jmine < len;
                double xold = 0.0, xbar = 0.0;
                if (mine && !raw) {   // in flight during the dot products
                    xold = P.Xv[p0 + jmine];
                    xbar = P.Xbarv[p0 + jmine];
                }
                T v[GH];
#pragma unroll
                for (int i = 0; i < GH; ++i) v[i] = (T)0;
                if (c * GH < len) {   // a group whose row is finished only takes part in the shuffles
#pragma unroll
                    for (int i = 0; i < GH; ++i) {
                        const int j = min(c * GH + i, len - 1);
                        const unsigned bbase = st.rows_a + lds_u16(la + 2u * (unsigned)j) * rowb;
                        V fa, fb;
                        LD::vec(fa, bbase + ca);
                        LD::vec(fb, bbase + cb2);
                        T d = (T)0;
#pragma unroll
                        for (int q = 0; q < VEC; ++q) d = fma(fka.v[q], fa.v[q], d);
#pragma unroll
                        for (int q = 0; q < VEC; ++q) d = fma(fkb.v[q], fb.v[q], d);
                        v[i] = d;
                    }
                }
                // transposing butterfly: after the stage with distance h a lane holds h sums
#pragma unroll
                for (int h = GH / 2; h >= 1; h >>= 1) {
                    const bool up = (lane & h) != 0;
#pragma unroll
                    for (int i = 0; i < h; ++i) {
                        const T send = up ? v[i] : v[i + h];
                        const T keep = up ? v[i + h] : v[i];
                        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
                    }
                }
                if (mine) {
                    const int p = p0 + jmine;
                    double x;
                    if (p == pd) {
                        x = xdiag;
                    } else {
                        x = (double)v[0] * inv_tr;
                        rsum += x;
                    }
                    if (raw) {
                        P.graw[p] = x;
                    } else {
                        P.Xbarv[p] = xbar + xold;
                        P.Xv[p] = x;
                    }
                }
            }
#pragma unroll
            for (int o = GH / 2; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
            {
                const double right = __shfl_down_sync(0xffffffffu, rsum, GH);
                if (role == 1) rsum += right;
            }
            if (valid && role != 2 && lane == 0 && !raw) {
                P.r[k] = rsum;
                if (S.pmask) push_f64(team, S, P.r + k, rsum, S.pmask[k]);
            }
        }
    }
}

// ===========================================================================
// Sketch-column sharding: after the ranks all-reduced [graw | dsq] (partial dot products and
// partial ||F_k||^2 over each rank's columns), every rank completes the Gram phase the same
// way: tr = sum dsq / K; X = graw / tr (diagonal: dsq / tr); X_avgd += X_prev; r = row sums.
template <typename T, int G, class Team>
__device__ void phase_gram_finish(const Prob<T>& P, const Team& team, double* sh) {
    const PlanDev& g = P.g;
    const int K = g.n;
    double trp = 0.0;
    for (int k = team.rank() * NT + threadIdx.x; k < K; k += team.size() * NT) trp += P.dsq[k];
    trp = block_sum(trp, sh);
    if (threadIdx.x == 0) P.ptr[team.rank()] = trp;
    team.sync(P);
    const double tr = team_sum(P.ptr, 1, team.size(), sh) / K;
    const int lane = threadIdx.x & (G - 1), grp = threadIdx.x / G;
    constexpr int R = NT / G;
    const int ntiles = (K + R - 1) / R;
    for (int t = team.rank(); t < ntiles; t += team.size()) {   // block-uniform: shuffles below are convergent
        const int k = t * R + grp;
        double rsum = 0.0;
        if (k < K) {
            const int pd = g.dpos[k], p1 = g.rowptr[k + 1];
            for (int p = g.rowptr[k] + lane; p < p1; p += G) {
                const double x = (p == pd ? P.dsq[k] : P.graw[p]) / tr;
                if (p != pd) rsum += x;
                P.Xbarv[p] += P.Xv[p];
                P.Xv[p] = x;
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
        if (k < K && lane == 0) P.r[k] = rsum;
    }
}

// ===========================================================================
// controller pieces shared by the fused kernel and the stepwise host loop
struct TaylorState {
    int m_star;
    long long s;
    double c1, a1, mu;
};

template <class Team, typename T>
__device__ __forceinline__ void taylor_begin(const Prob<T>& P, const Team& team, TaylorState& ts) {
    ts.a1 = team.a1(P);
    ts.c1 = team.c1(P);
    ts.mu = *reinterpret_cast<const volatile double*>(&P.ctrl->mu);
    taylor_select(ts.a1, ts.m_star, ts.s);
}

template <typename T>
__device__ __forceinline__ void record_history(const Prob<T>& P, long long iter, const TaylorState& ts, int nterms) {
    const int h = (int)(iter % HIST);
    P.hist_m[h] = ts.m_star;
    P.hist_s[h] = (int)(ts.s < 0x7fffffff ? ts.s : 0x7fffffff);
    P.hist_nt[h] = nterms;
    P.hist_a1[h] = ts.a1;
    P.hist_mu[h] = ts.mu;
}

template <typename T>
__device__ __forceinline__ void stage_setup(const Prob<T>& P, unsigned char* dyn, Stage<T>& st) {
    const TileDev& tl = P.tl;
    const unsigned base = smem_u32(dyn);
    const unsigned rows_bytes = ((unsigned)tl.ucap * (unsigned)P.Dp * (unsigned)sizeof(T) + 15u) & ~15u;
    const unsigned vals_bytes = ((unsigned)(tl.nnzcap + 4) * (unsigned)sizeof(T) + 15u) & ~15u;
    st.bar_a = base;
    st.rows_a = base + 16u;
    st.vals_a = st.rows_a + rows_bytes;
    st.lcol_a = st.vals_a + vals_bytes;
    st.parity = 0u;
    st.va = st.vals_a;
    st.la = st.lcol_a;
    if (tl.enabled) {
        if (threadIdx.x == 0) mbar_init(reinterpret_cast<unsigned long long*>(dyn), 1);
        __syncthreads();
    }
}

// The whole MMW loop for one team: n_iters iterations, no host involvement.  The leader
// thread timestamps the phases (globaltimer) for the reference's per-phase logs and adds up
// what it spends in team barriers (dbg[4], ns): for a row-sharded solver that is the
// exchange + cross-GPU wait of this rank.
template <typename T, int G, class Team>
__device__ void run_iterations(const Prob<T>& P, const Team& team, int n_iters, unsigned char* dyn, double* sh,
                               int do_finish = 0) {
    Stage<T> st;
    stage_setup(P, dyn, st);
    const bool staged = P.tl.enabled != 0;
    Ctrl* ctrl = P.ctrl;
    long long terms = 0;
    const bool leader = team.rank() == 0 && threadIdx.x == 0;
    unsigned long long sync_ns = 0;
#define TIMED_SYNC(WHAT, SLOT)                                    \
    do {                                                          \
        const unsigned long long ts0_ = leader ? globaltimer_ns() : 0ull; \
        team.sync(P, WHAT, SLOT);                                 \
        if (leader) sync_ns += globaltimer_ns() - ts0_;           \
    } while (0)
    if (do_finish) {   // split mode: complete the previous iteration's Gram from the all-reduced buffer
        phase_gram_finish<T, G>(P, team, sh);
        TIMED_SYNC(SY_PLAIN, 0);
    }
    double shift = *reinterpret_cast<const volatile double*>(&ctrl->smax_shift);
    for (int it = 0; it < n_iters; ++it) {
        unsigned long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        if (leader) t0 = globaltimer_ns();
        phase_dual<T, G>(P, team, shift, sh);
        TIMED_SYNC(SY_DUAL, 0);
        shift = team.emax(P);   // this iteration's maximum is the next one's soft-max shift
        if (leader) t1 = globaltimer_ns();
        phase_loss<T, G>(P, team, it, sh);
        TIMED_SYNC(SY_LOSS, 0);
        if (leader) t2 = globaltimer_ns();
        TaylorState ts;
        taylor_begin(P, team, ts);
        T* bin = P.B0;
        T* bout = P.B1;
        int tcount = 0;
        double c1 = ts.c1, fn_last = ts.c1;
        for (long long si = 0; si < ts.s; ++si) {
            if (si > 0) {
                phase_copy<T, G>(P, team, bin);
                c1 = fn_last;
                TIMED_SYNC(SY_PLAIN, 0);
            }
            for (int j = 0; j < ts.m_star; ++j) {
                const int slot = tcount % 3;
                const double coeff = 1.0 / ((double)ts.s * (double)(j + 1));
                if (staged && G >= 8 && P.Dp == G * Vec<T>::N)
                    phase_term_staged2<T, G>(P, team, bin, bout, coeff, slot, sh, st);
                else if (staged)
                    phase_term_staged<T, G>(P, team, bin, bout, coeff, slot, sh, st);
                else
                    phase_term<T, G>(P, team, bin, bout, coeff, ts.mu, slot, sh);
                TIMED_SYNC(SY_TERM, slot);
                double c2;
                team.term_norms(P, slot, c2, fn_last);
                T* tmp = bin;
                bin = bout;
                bout = tmp;
                ++tcount;
                if (c1 + c2 <= P.tol * fn_last) break;
                c1 = c2;
            }
        }
        terms += tcount;
        if (leader) {
            t3 = globaltimer_ns();
            ctrl->emax_key = 0ull;   // everyone read it right after the dual barrier, at least one barrier ago
        }
        if (staged)
            if (G >= 8 && P.Dp == G * Vec<T>::N)
                phase_gram_staged2<T, G>(P, team, sh, st);
            else
                phase_gram_staged<T, G>(P, team, sh, st);
        else
            phase_gram<T, G>(P, team, sh);
        TIMED_SYNC(SY_PLAIN, 0);
        if (leader) {
            const long long iter = ctrl->iter + it;
            record_history(P, iter, ts, tcount);
            const unsigned long long t4 = globaltimer_ns();
            double* ht = P.hist_t + (size_t)(iter % HIST) * 4;
            ht[0] = (double)(t1 - t0) * 1e-3;
            ht[1] = (double)(t2 - t1) * 1e-3;
            ht[2] = (double)(t3 - t2) * 1e-3;
            ht[3] = (double)(t4 - t3) * 1e-3;
        }
    }
#undef TIMED_SYNC
    if (leader) {
        ctrl->iter += n_iters;
        ctrl->smax_shift = shift;
        ctrl->total_terms += terms;
        ctrl->dbg[4] += (long long)sync_ns;
        team.finish(P);
    }
}

}  // namespace sigsdp
